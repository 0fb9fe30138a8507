"""HF-hub processor surface (hf_example.py:7-18) on the GPU preprocessing:

  OmChatImageProcessor   omchat/hf/image_processing_omchat.py:569-733  images -> {"pixel_values" [b, max_patches, 3, S, S]
                         (zero-padded along the patch axis, :530-567), "num_patches" [b]}
  OmChatProcessor        omchat/hf/processing_omchat.py:142-246        (text, images) -> {"input_ids" [1, T], "images" [n, 3, S, S]}
                         with the ChatML context of make_context and one "<image>" + (n-1) "patch:<image>" tags per image

Same geometry (select_best_resolution, _get_patch_output_size :110-125), same Pillow bicubic resampling, same rescale /
normalise as the mm_utils path of omchat_b200/preprocess.py. One deliberate difference: the reference pads the resized image
symmetrically (`((paste_y, paste_y), (paste_x, paste_x))`, :450-464), which leaves the canvas one pixel short when the
padding is odd and makes the last patch ragged (it is then re-resized by `_preprocess`); here the canvas always has the full
grid resolution, like mm_utils.resize_and_pad_image (the path cli.py uses). For even paddings the pixels are identical.
The reference's text-only branch builds `BatchFeature(data={**tensor})` (a TypeError); here it returns {"input_ids": ids}.
"""
from __future__ import annotations

from typing import List, Optional, Sequence

import torch

from .preprocess import IMAGE_MEAN, IMAGE_STD, AnyResPreprocessor
from .prompt import make_context

DEFAULT_PINPOINTS = [[896, 448], [448, 896], [896, 896], [448, 1344], [1344, 448]]  # image_processing_omchat.py:195-199


class BatchFeature(dict):
    """Minimal stand-in for transformers.BatchFeature: attribute access and .to(device)."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def to(self, device):
        return BatchFeature({k: (v.to(device) if isinstance(v, torch.Tensor) else v) for k, v in self.items()})


class OmChatImageProcessor:
    model_input_names = ["pixel_values"]

    def __init__(self, image_grid_pinpoints: Optional[List] = None, size: int = 448, image_mean=IMAGE_MEAN, image_std=IMAGE_STD,
                 device="cuda", dtype=torch.float32):
        self.image_grid_pinpoints = image_grid_pinpoints if image_grid_pinpoints is not None else DEFAULT_PINPOINTS
        self.size = {"shortest_edge": size}
        self.crop_size = {"height": size, "width": size}
        self._pre_args = dict(crop=size, device=device, dtype=dtype, mean=image_mean, std=image_std)
        self._pre_obj = None

    @property
    def _pre(self) -> AnyResPreprocessor:
        """Built on first use: the processor object itself (tokenizer + geometry) loads on a host without a GPU, the pixel
        work does not - AnyResPreprocessor raises there (no CPU fallback)."""
        if self._pre_obj is None:
            self._pre_obj = AnyResPreprocessor(self.image_grid_pinpoints, **self._pre_args)
        return self._pre_obj

    def preprocess(self, images, return_tensors="pt", **unused) -> BatchFeature:
        if not isinstance(images, (list, tuple)):
            images = [images]
        crops = [self._pre(im) for im in images]
        num = [c.shape[0] for c in crops]
        mx = max(num)
        padded = torch.zeros(len(crops), mx, *crops[0].shape[1:], device=crops[0].device, dtype=crops[0].dtype)
        for i, c in enumerate(crops):
            padded[i, : c.shape[0]] = c
        return BatchFeature(pixel_values=padded, num_patches=torch.tensor(num))

    __call__ = preprocess


class OmChatProcessor:
    attributes = ["image_processor", "tokenizer"]

    def __init__(self, image_processor: OmChatImageProcessor = None, tokenizer=None, **kwargs):
        self.image_processor = image_processor
        self.tokenizer = tokenizer

    def __call__(self, text: str, images=None, return_tensors="pt", **unused) -> BatchFeature:
        """processing_omchat.py:221-246."""
        system = "You are a helpful assistant."
        if images is None:
            _, ids = make_context(self.tokenizer, text.replace("<image>", "").strip(), None, system)
            return BatchFeature(input_ids=torch.tensor([ids]))
        feats = self.image_processor(images, return_tensors=return_tensors)
        num = feats["num_patches"].tolist()
        per_image = [feats["pixel_values"][i, :n] for i, n in enumerate(num)]  # split_tensor :131-140
        if len(per_image) == 1:
            n = num[0]
            query = "<image>\n" + "\n".join(["patch:<image>"] * (n - 1)) + "\n" + text.replace("<image>", "").strip()
        else:
            texts = text.split("<image>")
            final = texts[0]
            for i, n in enumerate(num):
                final += "<image>\n" + "\n".join(["patch:<image>"] * (n - 1))
                if i + 1 < len(texts):
                    final += texts[i + 1]
            query = final.strip()
        _, ids = make_context(self.tokenizer, query, None, system)
        return BatchFeature(input_ids=torch.tensor([ids]), images=torch.cat(per_image, dim=0))

    def batch_decode(self, *args, **kwargs):
        return self.tokenizer.batch_decode(*args, **kwargs)

    def decode(self, *args, **kwargs):
        return self.tokenizer.decode(*args, **kwargs)
