"""InternViT-6B (and the lighter InternViT-300M) vision tower + mm_projector on the C-ABI kernels.

Mirrors InternVITVisionTower (omchat/model/multimodal_encoder/internVIT_encoder.py:10-56) and the projector of
omchat/model/multimodal_projector/builder.py:54-61. Per layer (intern_vit_6b/modeling_intern_vit.py:218-220):
    rmsnorm -> QKV GEMM -> full-width q/k RMSNorm (in place) -> flash attention -> proj GEMM (+bias, *ls1, +residual)
    rmsnorm -> fc1 GEMM (+bias, GELU) -> fc2 GEMM (+bias, *ls2, +residual)
All activations are bf16 [rows, C] with rows = crops * (patches + 1); the residual stream is updated in place.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .. import lib
from ..config import OmChatQwen2Config
from .weights import ProjW, VitW, fold_norm, pad_head_cols, pad_head_rows

ATTN_HEAD_DIM = 128       # the attention kernels' head_dim ...
ATTN_NATIVE_DIMS = (64, 128)  # ... and the ones the tcgen05 kernel is instantiated for; other widths run zero-padded to 128


class InternVITVisionTower:
    """Same surface the reference callers touch: .is_loaded, .load_model(), .image_processor, .hidden_size,
    .num_patches, __call__(images) -> [n, P, C] (feature_select 'patch' drops CLS, internVIT_encoder.py:35-43)."""

    def __init__(self, cfg: OmChatQwen2Config, weights: Optional[VitW]):
        self.cfg = cfg
        self.vc = cfg.vision_config
        self.w = weights
        self.is_loaded = weights is not None
        self.select_layer = cfg.mm_vision_select_layer
        self.select_feature = cfg.mm_vision_select_feature
        self.image_processor = None  # CPU preprocessing (CLIPImageProcessor) is outside the hot path
        self.max_crops_per_pass = 64
        self._cu_cache = {}
        # norm1 / norm2 folded into the GEMMs that follow them (0 = stand-alone RMSNorm kernels, the round-1 path); RMSNorm only:
        # the 'layer_norm' variant (InternViT-300M) keeps its stand-alone LayerNorm kernel
        self.fold_norms = os.environ.get("OMCHAT_B200_FOLD_NORMS", "1") != "0" and self.vc.norm_type == "rms_norm"
        self._mats = {}
        self._ssq_bufs = None
        D = self.vc.head_dim
        if D > ATTN_HEAD_DIM or ATTN_HEAD_DIM % D != 0 or D % 8 != 0:
            raise ValueError(f"vision head_dim {D} not supported (must divide {ATTN_HEAD_DIM})")
        # head width the attention kernel runs at: D itself when there is an instantiation for it (64: InternViT-300M), else 128
        # with every head zero-padded (pad_heads = True forces that path; the tests keep it alive)
        self.pad_heads = D not in ATTN_NATIVE_DIMS
        if self.attn_dim != D and self.vc.qk_normalization:
            raise NotImplementedError("qk_normalization over zero-padded heads")

    @property
    def attn_dim(self) -> int:
        return ATTN_HEAD_DIM if self.pad_heads else self.vc.head_dim

    def load_model(self):
        if self.w is None:
            raise RuntimeError("vision tower weights were not provided")
        self.is_loaded = True

    @property
    def hidden_size(self) -> int:
        return self.vc.hidden_size

    @property
    def num_patches(self) -> int:
        return self.vc.num_patches

    def num_layers_to_run(self) -> int:
        L = self.vc.num_hidden_layers
        k = self.select_layer if self.select_layer >= 0 else L + 1 + self.select_layer
        if not 0 <= k <= L:
            raise ValueError(f"mm_vision_select_layer {self.select_layer} out of range")
        return k

    def _cu_seqlens(self, n: int, S: int, device) -> torch.Tensor:
        key = (n, S, str(device))
        if key not in self._cu_cache:
            self._cu_cache[key] = (torch.arange(n + 1, dtype=torch.int32) * S).to(device)
        return self._cu_cache[key]

    @torch.no_grad()
    def hidden_states(self, pixels: torch.Tensor, n_layers: Optional[int] = None, collect: bool = False):
        """Run embeddings + the first n_layers blocks. Returns hidden [n*(P+1), C] (and the per-layer list if collect)."""
        w, vc = self.w, self.vc
        if pixels.dim() != 4 or pixels.shape[1] != 3:
            raise ValueError(f"wrong pixel_values size: {tuple(pixels.shape)}")  # modeling_intern_vit.py:338
        if pixels.shape[2] != vc.image_size or pixels.shape[3] != vc.image_size:
            raise ValueError(f"expected {vc.image_size}x{vc.image_size} crops, got {tuple(pixels.shape[2:])}")
        n = pixels.shape[0]
        if pixels.dtype not in (torch.float32, torch.bfloat16):
            pixels = pixels.float()
        pixels = pixels.contiguous()
        C, H = vc.hidden_size, vc.num_attention_heads
        S = vc.num_patches + 1
        n_layers = self.num_layers_to_run() if n_layers is None else n_layers
        cols = lib.vit_im2col(pixels, w.patch_w.shape[1])
        patch = lib.gemm(cols, w.patch_w, bias=w.patch_b)
        h = lib.vit_assemble(patch, w.cls, w.pos, n)
        del cols, patch
        states = [h.clone()] if collect else None
        cu = self._cu_seqlens(n, S, h.device)
        rows = n * S
        Da = self.attn_dim
        Ca = H * Da  # attention width: = C unless the heads run zero-padded to 128 dims
        xn = torch.empty(rows, C, device=h.device, dtype=torch.bfloat16)
        qkv = torch.empty(rows, 3 * Ca, device=h.device, dtype=torch.bfloat16)
        attn = torch.empty(rows, Ca, device=h.device, dtype=torch.bfloat16)
        act = torch.empty(rows, vc.intermediate_size, device=h.device, dtype=torch.bfloat16)
        scale = (C // H) ** -0.5
        eps = vc.layer_norm_eps
        mats = self._layer_mats()
        if self.fold_norms:
            # norm1 / norm2 folded into the qkv / fc1 GEMMs (omc_gemm_bf16_norm): the residual epilogues of proj / fc2 leave
            # the rows' sums of squares behind, the next GEMM scales its rows by rstd - no stand-alone RMSNorm pass
            ssq_a, ssq_b = self._ssq(rows, h.device)
            ssq_a.from_rows(h)
            for li in range(n_layers):
                l, (qkv_f, qkv_b, proj_w, fc1_f) = w.layers[li], mats[li]
                lib.gemm(h, qkv_f, out=qkv, bias=qkv_b, ssq_in=ssq_a, norm_dim=C, eps=eps)
                if vc.qk_normalization:
                    lib.rmsnorm_pair(qkv, l.q_norm, l.k_norm, C, eps)
                lib.attention(qkv[:, :Ca], qkv[:, Ca:2 * Ca], qkv[:, 2 * Ca:], attn, cu, S, H, H, False, scale, head_dim=Da)
                lib.gemm(attn, proj_w, out=h, bias=l.proj_b, scale=l.ls1, res=h, epi=lib.EPI_RES, ssq_out=ssq_b)
                lib.gemm(h, fc1_f, out=act, bias=l.fc1_b, epi=lib.EPI_GELU, ssq_in=ssq_b, norm_dim=C, eps=eps)
                lib.gemm(act, l.fc2_w, out=h, bias=l.fc2_b, scale=l.ls2, res=h, epi=lib.EPI_RES, ssq_out=ssq_a)
                if collect:
                    states.append(h.clone())
            return (h, states) if collect else h
        layer_norm = vc.norm_type == "layer_norm"
        for li in range(n_layers):
            l, (qkv_w, qkv_b, proj_w, fc1_w) = w.layers[li], mats[li]
            if layer_norm:
                lib.layernorm(h, l.norm1, l.norm1_b, eps, out=xn)
            else:
                lib.rmsnorm(h, l.norm1, eps, out=xn)
            lib.gemm(xn, qkv_w, out=qkv, bias=qkv_b)
            if vc.qk_normalization:
                lib.rmsnorm(qkv[:, :C], l.q_norm, eps, out=qkv[:, :C])
                lib.rmsnorm(qkv[:, C:2 * C], l.k_norm, eps, out=qkv[:, C:2 * C])
            lib.attention(qkv[:, :Ca], qkv[:, Ca:2 * Ca], qkv[:, 2 * Ca:], attn, cu, S, H, H, False, scale, head_dim=Da)
            lib.gemm(attn, proj_w, out=h, bias=l.proj_b, scale=l.ls1, res=h, epi=lib.EPI_RES)
            if layer_norm:
                lib.layernorm(h, l.norm2, l.norm2_b, eps, out=xn)
            else:
                lib.rmsnorm(h, l.norm2, eps, out=xn)
            lib.gemm(xn, fc1_w, out=act, bias=l.fc1_b, epi=lib.EPI_GELU)
            lib.gemm(act, l.fc2_w, out=h, bias=l.fc2_b, scale=l.ls2, res=h, epi=lib.EPI_RES)
            if collect:
                states.append(h.clone())
        return (h, states) if collect else h

    def _layer_mats(self):
        """Per layer (qkv_w, qkv_b, proj_w, fc1_w) as the GEMMs take them, built once: norm1 / norm2 folded into qkv / fc1 when
        fold_norms (6.5 GB more for InternViT-6B), heads zero-padded to the attention kernels' 128 dims when narrower."""
        key = (self.fold_norms, self.pad_heads)
        if key not in self._mats:
            vc = self.vc
            H, D = vc.num_attention_heads, vc.head_dim
            mats = []
            for l in self.w.layers:
                qkv_w = fold_norm(l.qkv, l.norm1) if self.fold_norms else l.qkv
                fc1_w = fold_norm(l.fc1_w, l.norm2) if self.fold_norms else l.fc1_w
                qkv_b, proj_w = l.qkv_b, l.proj_w
                if self.attn_dim != D:
                    qkv_w = pad_head_rows(qkv_w, 3, H, D, ATTN_HEAD_DIM)
                    qkv_b = pad_head_rows(qkv_b, 3, H, D, ATTN_HEAD_DIM) if qkv_b is not None else None
                    proj_w = pad_head_cols(proj_w, H, D, ATTN_HEAD_DIM)
                mats.append((qkv_w, qkv_b, proj_w, fc1_w))
            self._mats[key] = mats
        return self._mats[key]

    def _folded(self):
        """(qkv * norm1, fc1 * norm2) per layer for the model-level C entry (lib.VitForward)."""
        return [(m[0], m[3]) for m in self._layer_mats()]

    def _ssq(self, rows: int, device):
        if self._ssq_bufs is None or self._ssq_bufs[0].rows < rows:
            self._ssq_bufs = (lib.RowSsq(rows, device), lib.RowSsq(rows, device))
        return self._ssq_bufs

    @torch.no_grad()
    def __call__(self, images: torch.Tensor, pixel_shuffle_down: int = 1) -> torch.Tensor:
        """images [n,3,S,S] -> features [n, L, C*down^2]; CLS dropped ('patch') then optional pixel shuffle."""
        if isinstance(images, (list, tuple)):
            images = torch.stack([im for im in images])
        if self.select_feature not in ("patch",):
            raise ValueError(f"Unexpected select feature: {self.select_feature}")
        vc = self.vc
        G = vc.image_size // vc.patch_size
        outs = []
        for i in range(0, images.shape[0], self.max_crops_per_pass):
            chunk = images[i:i + self.max_crops_per_pass]
            h = self.hidden_states(chunk)
            f = lib.select_pixel_shuffle(h, chunk.shape[0], G, pixel_shuffle_down)
            outs.append(f.view(chunk.shape[0], (G // pixel_shuffle_down) ** 2, -1))
        return outs[0] if len(outs) == 1 else torch.cat(outs, dim=0)

    forward = __call__


class InternVIT300mVisionTower(InternVITVisionTower):
    """The lighter tower (multimodal_encoder/internVIT300m_encoder.py:10-56, intern_vit_300m/modeling_intern_vit.py): LayerNorm
    instead of RMSNorm (norm_type = 'layer_norm', :61-64,209-210), no QK-norm, 16 heads of 64 dims, 24 layers. Same kernels, the
    attention kernel in its head_dim-64 instantiation."""

    def __init__(self, cfg: OmChatQwen2Config, weights: Optional[VitW]):
        super().__init__(cfg, weights)
        if self.vc.norm_type != "layer_norm" and self.vc.head_dim == ATTN_HEAD_DIM:
            raise ValueError("InternVIT300mVisionTower expects the 300M vision_config (InternVisionConfig.intern_vit_300m())")


def build_vision_tower(cfg: OmChatQwen2Config, weights: Optional[VitW]) -> InternVITVisionTower:
    """multimodal_encoder/builder.py:7-18: the tower class follows the NAME in mm_vision_tower."""
    name = (cfg.mm_vision_tower or "").lower()
    if "internvit-300m" in name:
        return InternVIT300mVisionTower(cfg, weights)
    return InternVITVisionTower(cfg, weights)


class MMProjector:
    """mlp2x_gelu projector (multimodal_projector/builder.py:54-61): Linear + GELU + Linear on [n*L, Cin]."""

    def __init__(self, weights: ProjW):
        self.w = weights

    @torch.no_grad()
    def __call__(self, feats: torch.Tensor) -> torch.Tensor:
        shp = feats.shape
        x = feats.reshape(-1, shp[-1])
        h = lib.gemm(x, self.w.w0, bias=self.w.b0, epi=lib.EPI_GELU)
        y = lib.gemm(h, self.w.w2, bias=self.w.b2)
        return y.view(*shp[:-1], y.shape[-1])
