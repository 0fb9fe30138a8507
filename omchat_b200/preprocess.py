"""GPU any-resolution preprocessing — the host side of csrc/preprocess.cu, mirroring the reference's helpers:

  select_best_resolution        omchat/mm_utils.py:12-40
  get_anyres_image_grid_shape   omchat/mm_utils.py:97-116
  process_anyres_image          omchat/mm_utils.py:119-158  -> AnyResPreprocessor.__call__ (same crop order: whole image
                                resized to crop x crop first, then the canvas patches row-major)
  process_images ("anyres")     omchat/mm_utils.py:164-182  -> AnyResPreprocessor.process_images

The resampling weights are Pillow's (src/libImaging/Resample.c precompute_coeffs + normalize_coeffs_8bpc), computed here in
float64 with the same operation order (sequential weight sum, division, round-half-away 22-bit fixed point) and applied by
omc_resample_u8 in exact integer arithmetic, so the uint8 pixels equal Image.resize()'s bit for bit. There is no CPU
fallback: the kernels need a CUDA device.
"""
from __future__ import annotations

import ast
import math
from typing import List, Sequence, Tuple, Union

import numpy as np
import torch

from . import lib

PRECISION_BITS = 32 - 8 - 2
IMAGE_MEAN = (0.485, 0.456, 0.406)  # internVIT_encoder.py:28
IMAGE_STD = (0.229, 0.224, 0.225)


def select_best_resolution(original_size: Tuple[int, int], possible_resolutions: Sequence[Sequence[int]]) -> Tuple[int, int]:
    """mm_utils.py:12-40: the grid resolution (width, height) that keeps most of the image and wastes least canvas."""
    original_width, original_height = original_size
    best_fit, max_effective, min_wasted = None, 0, float("inf")
    for width, height in possible_resolutions:
        scale = min(width / original_width, height / original_height)
        down_w, down_h = int(original_width * scale), int(original_height * scale)
        effective = min(down_w * down_h, original_width * original_height)
        wasted = width * height - effective
        if effective > max_effective or (effective == max_effective and wasted < min_wasted):
            max_effective, min_wasted, best_fit = effective, wasted, (width, height)
    return best_fit


def _pinpoints(grid_pinpoints) -> List[Sequence[int]]:
    return grid_pinpoints if isinstance(grid_pinpoints, list) else ast.literal_eval(grid_pinpoints)


def get_anyres_image_grid_shape(image_size: Tuple[int, int], grid_pinpoints, patch_size: int) -> Tuple[int, int]:
    """mm_utils.py:97-116."""
    width, height = select_best_resolution(image_size, _pinpoints(grid_pinpoints))
    return width // patch_size, height // patch_size


def resize_and_pad_geometry(original_size: Tuple[int, int], target_resolution: Tuple[int, int]):
    """mm_utils.py:54-71: (new_w, new_h, paste_x, paste_y) of the aspect-preserving resize on the black canvas."""
    ow, oh = original_size
    tw, th = target_resolution
    scale_w, scale_h = tw / ow, th / oh
    if scale_w < scale_h:
        new_w, new_h = tw, min(math.ceil(oh * scale_w), th)
    else:
        new_h, new_w = th, min(math.ceil(ow * scale_h), tw)
    return new_w, new_h, (tw - new_w) // 2, (th - new_h) // 2


def _bicubic(x: np.ndarray, a: float = -0.5) -> np.ndarray:
    x = np.abs(x)
    near = ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    far = (((x - 5) * x + 8) * x - 4) * a
    return np.where(x < 1.0, near, np.where(x < 2.0, far, 0.0))


def precompute_coeffs(in_size: int, out_size: int):
    """Pillow Resample.c precompute_coeffs (bicubic, whole-image box) + normalize_coeffs_8bpc, vectorised with the same
    float64 operation order. Returns (coefs int32 [out, ksize], bounds int32 [out, 2] = (first tap, number of taps))."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    ss = 1.0 / filterscale
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast: truncation toward zero
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size) - xmin
    k = np.arange(ksize, dtype=np.int64)[None, :]
    w = _bicubic(((k + xmin[:, None]) - center[:, None] + 0.5) * ss)
    w = np.where(k < xmax[:, None], w, 0.0)
    ww = np.cumsum(w, axis=1)[:, -1:]                                          # sequential sum, like the C loop
    kk = np.where(ww != 0.0, w / np.where(ww != 0.0, ww, 1.0), w)
    fixed = np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))
    coefs = np.trunc(fixed).astype(np.int32)
    bounds = np.stack([xmin, xmax], axis=1).astype(np.int32)
    return np.ascontiguousarray(coefs), np.ascontiguousarray(bounds)


def normalize_lut(mean=IMAGE_MEAN, std=IMAGE_STD, rescale: float = 1 / 255) -> np.ndarray:
    """[3, 256] fp32 table of CLIPImageProcessor's rescale + normalize (transformers 4.41 image_transforms: value * scale
    in float64 cast to float32, then (x - mean) / std in float32)."""
    v = (np.arange(256, dtype=np.float64) * rescale).astype(np.float32)
    m = np.asarray(mean, dtype=np.float32)[:, None]
    s = np.asarray(std, dtype=np.float32)[:, None]
    return ((v[None, :] - m) / s).astype(np.float32)


class AnyResPreprocessor:
    """process_anyres_image on the GPU. __call__(image) takes an RGB uint8 image as a PIL.Image, a numpy array [H, W, 3] or
    a torch uint8 tensor [H, W, 3] (host or device) and returns the crops [1 + patches, 3, crop, crop] on `device`."""

    def __init__(self, grid_pinpoints, crop: int = 448, device="cuda", dtype=torch.float32, mean=IMAGE_MEAN, std=IMAGE_STD):
        if not torch.cuda.is_available():
            raise lib.OmcError("omchat_b200 preprocessing needs a CUDA device (no CPU fallback)")
        lib.load()
        self.grid_pinpoints = _pinpoints(grid_pinpoints)
        self.crop = crop
        self.device = torch.device(device)
        self.dtype = dtype
        self.lut = torch.from_numpy(normalize_lut(mean, std)).to(self.device)
        self._coefs = {}

    def _tables(self, in_size: int, out_size: int):
        key = (in_size, out_size)
        t = self._coefs.get(key)
        if t is None:
            if len(self._coefs) > 256:
                self._coefs.clear()
            c, b = precompute_coeffs(in_size, out_size)
            t = (torch.from_numpy(c).to(self.device), torch.from_numpy(b).to(self.device))
            self._coefs[key] = t
        return t

    def resize(self, img: torch.Tensor, size: Tuple[int, int]) -> torch.Tensor:
        """PIL Image.resize((w, h)) (BICUBIC) of a device uint8 image [H, W, 3]: horizontal pass, then vertical."""
        w, h = size
        H, W = img.shape[0], img.shape[1]
        if (W, H) == (w, h):
            return img
        t = img
        if W != w:
            c, b = self._tables(W, w)
            t = lib.resample_u8(t, H, w, c, b, vertical=False)
        if H != h:
            c, b = self._tables(H, h)
            t = lib.resample_u8(t, h, w, c, b, vertical=True)
        return t

    def _to_device_u8(self, image) -> torch.Tensor:
        if isinstance(image, torch.Tensor):
            t = image
        elif isinstance(image, np.ndarray):
            t = torch.from_numpy(np.ascontiguousarray(image))
        else:  # PIL.Image (process_anyres_image's input type); convert like CLIPImageProcessor's do_convert_rgb
            t = torch.from_numpy(np.asarray(image.convert("RGB")).copy())
        if t.dtype != torch.uint8 or t.dim() != 3 or t.shape[2] != 3:
            raise ValueError(f"expected an RGB uint8 image [H, W, 3], got {tuple(t.shape)} {t.dtype}")
        return t.to(self.device, non_blocking=True).contiguous()

    @torch.no_grad()
    def __call__(self, image, return_best_res: bool = False):
        img = self._to_device_u8(image)
        H, W = img.shape[0], img.shape[1]
        best = select_best_resolution((W, H), self.grid_pinpoints)
        new_w, new_h, px, py = resize_and_pad_geometry((W, H), best)
        resized = self.resize(img, (new_w, new_h))
        thumb = self.resize(img, (self.crop, self.crop))
        out = lib.anyres_pack(thumb, resized, best[0], best[1], px, py, self.crop, self.lut, self.dtype)
        return (out, best) if return_best_res else out

    def process_images(self, images) -> Union[torch.Tensor, List[torch.Tensor]]:
        """mm_utils.py:164-182 for image_aspect_ratio == 'anyres': a stacked tensor when every image yields the same number
        of crops, else a list."""
        outs = [self(im) for im in images]
        if all(o.shape == outs[0].shape for o in outs):
            return torch.stack(outs, dim=0)
        return outs
