"""CPU restatement (numpy) of the reference's any-resolution image preprocessing — TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench tooling may import this module; the product path (omchat_b200/preprocess.py)
never does. What is restated:
  * select_best_resolution                         omchat/mm_utils.py:12-40
  * resize_and_pad_image                           omchat/mm_utils.py:43-73   (PIL Image.resize default = BICUBIC, black canvas)
  * divide_to_patches                              omchat/mm_utils.py:76-94
  * process_anyres_image                           omchat/mm_utils.py:119-158 (thumbnail first, then the canvas patches)
  * CLIPImageProcessor.preprocess on a crop that already has the target size (internVIT_encoder.py:26-29: size = crop_size
    = 448, ImageNet mean/std): resize and centre crop are identities, then rescale by 1/255 and normalise
    (transformers image_processing_clip / image_transforms.rescale + normalize)
  * the resampling itself lives in a third-party dependency, Pillow (pin: none in pyproject.toml; checked copy 12.2.0):
    src/libImaging/Resample.c — precompute_coeffs (support scaled by the down-scale factor, normalised double weights),
    normalize_coeffs_8bpc (22-bit fixed point, round half away from zero), ImagingResampleHorizontal_8bpc /
    ImagingResampleVertical_8bpc (accumulator seeded with 1 << 21, >> 22, clip to 0..255), horizontal pass first with a
    uint8 intermediate. bicubic_filter with a = -0.5.
Pinned: tests/test_preprocess.py checks `resize` bit-for-bit against Pillow itself on random images (up- and down-scaling,
odd sizes) and `process_anyres` against vectors produced by the real reference function
(tests/golden/make_golden_preprocess.py -> tests/golden/golden_preprocess.npz).
"""
from __future__ import annotations

import math
from typing import List, Sequence, Tuple

import numpy as np

PRECISION_BITS = 32 - 8 - 2
IMAGE_MEAN = (0.485, 0.456, 0.406)
IMAGE_STD = (0.229, 0.224, 0.225)


def select_best_resolution(original_size: Tuple[int, int], possible_resolutions: Sequence[Sequence[int]]) -> Tuple[int, int]:
    """mm_utils.py:12-40 — (width, height) in, best (width, height) out."""
    ow, oh = original_size
    best, max_eff, min_waste = None, 0, float("inf")
    for w, h in possible_resolutions:
        scale = min(w / ow, h / oh)
        dw, dh = int(ow * scale), int(oh * scale)
        eff = min(dw * dh, ow * oh)
        waste = w * h - eff
        if eff > max_eff or (eff == max_eff and waste < min_waste):
            max_eff, min_waste, best = eff, waste, (w, h)
    return best


def _bicubic(x: float, a: float = -0.5) -> float:
    x = abs(x)
    if x < 1.0:
        return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
    if x < 2.0:
        return (((x - 5) * x + 8) * x - 4) * a
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c precompute_coeffs + normalize_coeffs_8bpc for the whole-image box. Returns (int32 [out, ksize], bounds)."""
    scale = in_size / out_size
    filterscale = max(scale, 1.0)
    support = 2.0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    kk = np.zeros((out_size, ksize), dtype=np.float64)
    bounds = np.zeros((out_size, 2), dtype=np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = max(int(center - support + 0.5), 0)
        xmax = min(int(center + support + 0.5), in_size) - xmin
        ww = 0.0
        for x in range(xmax):
            w = _bicubic((x + xmin - center + 0.5) * ss)
            kk[xx, x] = w
            ww += w
        if ww != 0.0:
            kk[xx, :xmax] /= ww
        bounds[xx] = (xmin, xmax)
    ki = np.trunc(np.where(kk < 0, -0.5 + kk * (1 << PRECISION_BITS), 0.5 + kk * (1 << PRECISION_BITS))).astype(np.int32)
    return ki, bounds


def _resample_axis1(img: np.ndarray, out_w: int) -> np.ndarray:
    H, W, C = img.shape
    ki, b = precompute_coeffs(W, out_w)
    out = np.zeros((H, out_w, C), dtype=np.uint8)
    for xx in range(out_w):
        xmin, n = int(b[xx, 0]), int(b[xx, 1])
        acc = np.full((H, C), 1 << (PRECISION_BITS - 1), dtype=np.int64)
        acc += (img[:, xmin:xmin + n, :].astype(np.int64) * ki[xx, :n][None, :, None]).sum(1)
        out[:, xx, :] = np.clip(acc >> PRECISION_BITS, 0, 255)
    return out


def resize(img: np.ndarray, size: Tuple[int, int]) -> np.ndarray:
    """PIL.Image.resize((w, h)) for an RGB uint8 image [H, W, 3] (default resample = BICUBIC, whole-image box)."""
    w, h = size
    H, W, _ = img.shape
    if (W, H) == (w, h):
        return img.copy()
    t = _resample_axis1(img, w) if W != w else img
    if H != h:
        t = _resample_axis1(np.ascontiguousarray(t.transpose(1, 0, 2)), h).transpose(1, 0, 2)
    return np.ascontiguousarray(t)


def resize_and_pad_geometry(original_size: Tuple[int, int], target_resolution: Tuple[int, int]):
    """mm_utils.py:54-71 — (new_w, new_h, paste_x, paste_y)."""
    ow, oh = original_size
    tw, th = target_resolution
    sw, sh = tw / ow, th / oh
    if sw < sh:
        nw, nh = tw, min(math.ceil(oh * sw), th)
    else:
        nh, nw = th, min(math.ceil(ow * sh), tw)
    return nw, nh, (tw - nw) // 2, (th - nh) // 2


def resize_and_pad(img: np.ndarray, target_resolution: Tuple[int, int]) -> np.ndarray:
    H, W, _ = img.shape
    nw, nh, px, py = resize_and_pad_geometry((W, H), target_resolution)
    canvas = np.zeros((target_resolution[1], target_resolution[0], 3), dtype=np.uint8)
    canvas[py:py + nh, px:px + nw] = resize(img, (nw, nh))
    return canvas


def normalize_lut() -> np.ndarray:
    """[3, 256] float32: CLIPImageProcessor rescale (x * 1/255, computed in float64 then cast to float32) + normalise
    ((x - mean) / std in float32) for every possible pixel value of every channel."""
    v = (np.arange(256, dtype=np.float64) * (1 / 255)).astype(np.float32)
    mean = np.asarray(IMAGE_MEAN, dtype=np.float32)[:, None]
    std = np.asarray(IMAGE_STD, dtype=np.float32)[:, None]
    return ((v[None, :] - mean) / std).astype(np.float32)


def process_anyres(img: np.ndarray, grid_pinpoints: Sequence[Sequence[int]], crop: int = 448) -> np.ndarray:
    """mm_utils.py:119-158 with the CLIP processor of internVIT_encoder.py:26-29. img uint8 [H, W, 3] ->
    float32 [1 + n_patches, 3, crop, crop]: the resized whole image first, then the canvas patches in row-major order."""
    H, W, _ = img.shape
    best = select_best_resolution((W, H), grid_pinpoints)
    canvas = resize_and_pad(img, best)
    crops = [resize(img, (crop, crop))]
    for i in range(0, best[1], crop):
        for j in range(0, best[0], crop):
            crops.append(canvas[i:i + crop, j:j + crop])
    lut = normalize_lut()
    out = np.empty((len(crops), 3, crop, crop), dtype=np.float32)
    for n, c in enumerate(crops):
        for ch in range(3):
            out[n, ch] = lut[ch][c[:, :, ch]]
    return out
