"""ctypes binding of libomchat_b200.so (the C-ABI declared in include/omchat_b200.h) + thin torch-tensor wrappers.

There is no fallback: if the shared library is missing or a call fails, an exception is raised.
"""
from __future__ import annotations

import ctypes
import os
from ctypes import c_char_p, c_float, c_int, c_longlong, c_void_p
from pathlib import Path
from typing import Optional

import torch

_PKG = Path(__file__).resolve().parent
LIB_PATH = _PKG / "_lib" / "libomchat_b200.so"

EPI_NONE, EPI_GELU, EPI_RES, EPI_SWIGLU = 0, 1, 2, 3

# every symbol include/omchat_b200.h declares: name -> (restype, argtypes)
_P, _I, _L, _F = c_void_p, c_int, c_longlong, c_float
SIGNATURES = {
    "omc_last_error": (c_char_p, []),
    "omc_version": (_I, []),
    "omc_num_sms": (_I, []),
    "omc_gemm_bf16": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _L, _I, _I, _I, _P]),
    "omc_gemm_bf16_norm": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _L, _I, _I, _I, _P, _P]),
    "omc_row_ssq_rows": (_I, [_P, _L, _L, _I, _P, _P]),
    "omc_gemm_skinny_workspace_bytes": (_L, [_I]),
    "omc_gemm_skinny_bf16": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _P, _P, _L, _I, _I, _P, _L, _P]),
    "omc_packed_weight_bytes": (_L, [_I, _I]),
    "omc_pack_weight": (_I, [_P, _L, _I, _I, _P, _P, _P]),
    "omc_gemm_stream_workspace_bytes": (_L, []),
    "omc_gemm_stream": (_I, [_P, _L, _I, _P, _I, _I, _P, _L, _I, _P, _P, _L, _I, _P, _I, _I, _F, _P, _P, _I, _P, _P]),
    "omc_gemm_stream_xchg_bytes": (_L, []),
    "omc_row_ssq": (_I, [_P, _L, _I, _I, _P, _I, _I, _P]),
    "omc_gemm_stream_set_prof": (_I, [_P, _I]),
    "omc_gemm_stream_set_debug": (_I, [_P]),
    "omc_gemv_bf16": (_I, [_P, _L, _P, _L, _P, _L, _I, _I, _I, _P, _F, _P, _P, _L, _I, _I, _P]),
    "omc_rmsnorm": (_I, [_P, _L, _P, _P, _L, _I, _I, _F, _P]),
    "omc_rmsnorm_pair": (_I, [_P, _L, _P, _P, _I, _I, _F, _P]),
    "omc_layernorm": (_I, [_P, _L, _P, _P, _P, _L, _I, _I, _F, _P]),
    "omc_moe_max_tiles": (_I, [_I, _I, _I]),
    "omc_moe_set_pdl": (_I, [_I]),
    "omc_moe_route": (_I, [_P, _L, _I, _I, _P, _F, _P, _L, _P, _P, _I, _I, _I, _P, _P, _P, _P, _P]),
    "omc_moe_select": (_I, [_P, _L, _I, _I, _I, _I, _I, _P, _P, _P, _P, _P]),
    "omc_moe_plan": (_I, [_P, _I, _I, _P, _P, _P, _P]),
    "omc_moe_scatter": (_I, [_P, _L, _I, _I, _P, _I, _P, _P, _P, _L, _P, _P]),
    "omc_gemm_bf16_grouped": (_I, [_P, _L, _I, _P, _L, _I, _I, _I, _P, _I, _P, _L, _I, _P]),
    "omc_moe_combine": (_I, [_P, _L, _I, _I, _P, _L, _P, _P, _I, _P, _L, _P, _P, _I, _P]),
    "omc_moe_plan_scatter": (_I, [_P, _I, _I, _P, _P, _P, _P, _L, _I, _I, _P, _I, _P, _L, _P, _P]),
    "omc_vit_im2col": (_I, [_P, _I, _P, _L, _I, _I, _I, _P]),
    "omc_vit_assemble": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "omc_select_pixel_shuffle": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "omc_resample_u8": (_I, [_P, _I, _I, _P, _I, _I, _P, _P, _I, _I, _P]),
    "omc_anyres_pack": (_I, [_P, _P, _I, _I, _I, _I, _I, _I, _I, _P, _P, _I, _P]),
    "omc_attention_fwd": (_I, [_P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _L, _I, _I, _I, _F, _P]),
    "omc_attention_fwd_hd": (_I, [_P, _L, _P, _L, _P, _L, _P, _L, _P, _I, _I, _L, _I, _I, _I, _I, _F, _P]),
    "omc_attention_set_impl": (_I, [_I]),
    "omc_attention_set_prof": (_I, [_P]),
    "omc_rope_kv_store": (_I, [_P, _L, _P, _P, _I, _I, _I, _P, _P, _P, _I, _I, _P]),
    "omc_decode_attn_splits": (_I, [_I, _I, _I]),
    "omc_decode_attn_workspace_bytes": (_L, [_I, _I, _I, _I]),
    "omc_paged_decode_attn": (_I, [_P, _L, _P, _P, _P, _I, _I, _P, _I, _I, _I, _I, _F, _P, _L, _P, _P]),
    "omc_embed_lookup": (_I, [_P, _I, _P, _I, _P, _L, _I, _P]),
    "omc_splice": (_I, [_P, _P, _I, _I, _L, _P, _P, _I, _I, _I, _I, _P, _P, _P, _P, _P, _I, _I, _P]),
    "omc_argmax": (_I, [_P, _L, _I, _I, _P, _P, _P]),
    "omc_vit_workspace_bytes": (_L, [_P, _I]),
    "omc_vit_forward": (_I, [_P, _P, _I, _I, _P, _P, _P]),
    "omc_decoder_prefill_workspace_bytes": (_L, [_P, _I, _I]),
    "omc_decoder_prefill": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P, _P, _P, _I, _P]),
    "omc_decode_plan_bytes": (_L, [_I]),
    "omc_decode_workspace_bytes": (_L, [_P]),
    "omc_decode_plan_build": (_I, [_P, _P]),
    "omc_decode_step": (_I, [_P, _P, ctypes.c_uint, _P]),
    "omc_decode_xchg_bytes": (_L, [_P]),
    "omc_peer_alloc": (_I, [_L, ctypes.POINTER(c_void_p), _P]),
    "omc_peer_open": (_I, [_P, ctypes.POINTER(c_void_p)]),
    "omc_peer_close": (_I, [_P]),
    "omc_peer_free": (_I, [_P]),
}

_lib = None
_launches = 0  # kernels launched through this binding (bench.py's gpu_launches; graph replays are added by the caller)


def launch_count() -> int:
    return _launches


def add_launches(n: int):
    global _launches
    _launches += n


class OmcError(RuntimeError):
    pass


def load() -> ctypes.CDLL:
    """Load the shared library (building is __graft_entry__.build()'s job; a missing library is a hard error)."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("OMCHAT_B200_LIB", str(LIB_PATH))
    if not os.path.exists(path):
        raise OmcError(f"{path} not found: run `python -m omchat_b200.build` (there is no CPU fallback)")
    lib = ctypes.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


_KERNELS_PER_CALL = {"omc_splice": 2, "omc_argmax": 2}


def _check(rc: int, what: str):
    global _launches
    _launches += _KERNELS_PER_CALL.get(what, 1)
    if rc != 0:
        msg = load().omc_last_error().decode(errors="replace")
        raise OmcError(f"{what} failed ({rc}): {msg}")


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _need_cuda(*ts):
    """Every kernel is launched on torch.cuda.current_stream() of the CURRENT device: a tensor living on another GPU would
    be dereferenced there (illegal access or silent peer traffic), so that is an error, not a convention."""
    cur = None
    for t in ts:
        if t is None:
            continue
        if not t.is_cuda:
            raise OmcError("omchat_b200 kernels need CUDA tensors (no CPU fallback)")
        if cur is None:
            cur = torch.cuda.current_device()
        if t.device.index != cur:
            raise OmcError(f"tensor on cuda:{t.device.index} but the current device is cuda:{cur}: run the call under "
                           f"`with torch.cuda.device({t.device.index}):` (the model classes do this themselves)")


SKINNY_MAX_M = 64  # rows up to which gemm() streams the weights through the swapped-operand skinny kernel
SKINNY_ENABLED = os.environ.get("OMCHAT_B200_NO_SKINNY", "0") != "1"
_skinny_ws = {}


def _skinny_workspace(device, n: int) -> torch.Tensor:
    """Zeroed split-K workspace of the skinny GEMM, one per device, grown on demand (the kernel leaves it zeroed)."""
    key = str(device)
    ws = _skinny_ws.get(key)
    need = load().omc_gemm_skinny_workspace_bytes(max(n, 160 * 1024))
    if ws is None or ws.numel() < need:
        ws = torch.zeros(need, device=device, dtype=torch.uint8)
        _skinny_ws[key] = ws
    return ws


# ----------------------------------------------------------------------------------------------- wrappers
# ---- tile choice for mid-sized M (one crop, one prompt: M ~ 1000): with 5-9 tiles along M and a 148-SM device the
# number of WAVES decides, not the per-tile efficiency - the library default (2-CTA pairs, widest BN dividing N) is up to
# 35 % off the best configuration there (tools/bench_gemm.py --small). Every configuration accumulates along K in the same
# order, so the choice does not change the result bits; it is made once per (M, N, K, epilogue) by timing the six
# instantiations on the caller's operands (scratch output, L2 flushed between launches) and cached.
GEMM_AUTOTUNE = os.environ.get("OMCHAT_B200_GEMM_AUTOTUNE", "1") != "0"
GEMM_AUTOTUNE_MAX_M = 2048
_GEMM_CFGS = [256 | (2 << 16), 192 | (2 << 16), 160 | (2 << 16), 128 | (2 << 16), 256 | (1 << 16), 128 | (1 << 16)]
_gemm_tuned: dict = {}
_l2_flush = None


def _gemm_autotune(x, w, out, bias, scale, res, epi, out_f32) -> int:
    global _l2_flush
    M, K = x.shape
    N = w.shape[0]
    key = (M, N, K, epi, out_f32, bias is not None, scale is not None)
    cfg = _gemm_tuned.get(key)
    if cfg is not None:
        return cfg
    if torch.cuda.is_current_stream_capturing():
        return 0
    if _l2_flush is None or _l2_flush.device != x.device:
        _l2_flush = torch.empty(256 << 20, device=x.device, dtype=torch.uint8)
    scratch = torch.empty_like(out)
    cands = [c for c in _GEMM_CFGS if epi != EPI_SWIGLU or (c & 0xFFFF) == 256]
    best, best_t = 0, float("inf")
    for c in cands:
        ts = []
        for i in range(4):
            _l2_flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = load().omc_gemm_bf16(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(scratch), scratch.stride(0), M, N, K,
                                      _ptr(bias), _ptr(scale), _ptr(res), res.stride(0) if res is not None else 0, epi,
                                      1 if out_f32 else 0, c, _stream())
            e1.record()
            if rc != 0:
                break
            if i > 0:
                ts.append((e0, e1))
        else:
            torch.cuda.synchronize()
            t = sorted(a.elapsed_time(b) for a, b in ts)[len(ts) // 2]
            if t < best_t:
                best, best_t = c, t
    _gemm_tuned[key] = best
    return best


class GemmNorm(ctypes.Structure):
    """Mirror of `omc_gemm_norm` (include/omchat_b200.h)."""
    _fields_ = [("ssq_in", c_void_p), ("ssq_in_ld", c_longlong), ("ssq_in_parts", c_int), ("norm_dim", c_int), ("eps", c_float),
                ("ssq_out_max_parts", c_int), ("ssq_out", c_void_p), ("ssq_out_ld", c_longlong), ("ssq_out_parts", c_int),
                ("reserved", c_int)]


class RowSsq:
    """Per-row sums of squares of a residual stream, as per-N-tile partials [max_parts, rows] fp32 (+ how many parts the
    last producer wrote): what an EPI_RES GEMM leaves behind for the folded RMSNorm of the GEMM that reads the rows next."""
    MAX_PARTS = 32

    def __init__(self, rows: int, device):
        self.buf = torch.empty(self.MAX_PARTS, rows, device=device, dtype=torch.float32)
        self.rows, self.parts = rows, 0

    def from_rows(self, x: torch.Tensor):
        """Rows no GEMM produced (embeddings, all-reduced rows): one warp per row."""
        _need_cuda(x)
        assert x.dim() == 2 and x.stride(1) == 1 and x.shape[0] <= self.rows
        rc = load().omc_row_ssq_rows(_ptr(x), x.stride(0), x.shape[0], x.shape[1], self.buf.data_ptr(), _stream())
        _check(rc, "omc_row_ssq_rows")
        self.parts = 1
        return self


def gemm(x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None, *, bias=None, scale=None, res=None,
         epi: int = EPI_NONE, out_f32: bool = False, tile_cfg: int = 0, ssq_in: Optional[RowSsq] = None, norm_dim: int = 0,
         eps: float = 1e-6, ssq_out: Optional[RowSsq] = None) -> torch.Tensor:
    """out[M,N] = epi(x[M,K] @ w[N,K]^T). x/out may be row-strided 2-D views (last dim contiguous).
    ssq_in: the RMSNorm in front of this layer is folded into w's columns; rows are scaled by rsqrt(sum / norm_dim + eps) in
    the epilogue (omc_gemm_bf16_norm). ssq_out: leave the sums of squares of the rows written for the next such GEMM."""
    _need_cuda(x, w)
    assert x.dim() == 2 and w.dim() == 2 and x.shape[1] == w.shape[1] and x.stride(1) == 1 and w.stride(1) == 1
    M, K = x.shape
    N = w.shape[0]
    n_out = N // 2 if epi == EPI_SWIGLU else N
    if out is None:
        out = torch.empty(M, n_out, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    if ssq_in is not None or ssq_out is not None:
        if tile_cfg == 0 and GEMM_AUTOTUNE and M <= GEMM_AUTOTUNE_MAX_M:
            tile_cfg = _gemm_autotune(x, w, out, bias, scale, res, epi, out_f32)
        nf = GemmNorm()
        if ssq_in is not None:
            assert ssq_in.parts >= 1 and ssq_in.rows >= M and norm_dim > 0
            nf.ssq_in, nf.ssq_in_ld, nf.ssq_in_parts, nf.norm_dim, nf.eps = ssq_in.buf.data_ptr(), ssq_in.rows, ssq_in.parts, norm_dim, eps
        if ssq_out is not None:
            assert ssq_out.rows >= M
            nf.ssq_out, nf.ssq_out_ld, nf.ssq_out_max_parts = ssq_out.buf.data_ptr(), ssq_out.rows, RowSsq.MAX_PARTS
        rc = load().omc_gemm_bf16_norm(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), M, N, K,
                                       _ptr(bias), _ptr(scale), _ptr(res), res.stride(0) if res is not None else 0, epi,
                                       1 if out_f32 else 0, tile_cfg, ctypes.byref(nf), _stream())
        _check(rc, "omc_gemm_bf16_norm")
        if ssq_out is not None:
            ssq_out.parts = nf.ssq_out_parts
        return out
    if M <= SKINNY_MAX_M and tile_cfg == 0 and SKINNY_ENABLED:
        ws = _skinny_workspace(x.device, N)
        rc = load().omc_gemm_skinny_bf16(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), M, N, K,
                                         _ptr(bias), _ptr(scale), _ptr(res), res.stride(0) if res is not None else 0, epi,
                                         1 if out_f32 else 0, ws.data_ptr(), ws.numel(), _stream())
        _check(rc, "omc_gemm_skinny_bf16")
        return out
    if tile_cfg == 0 and GEMM_AUTOTUNE and M <= GEMM_AUTOTUNE_MAX_M:
        tile_cfg = _gemm_autotune(x, w, out, bias, scale, res, epi, out_f32)
    rc = load().omc_gemm_bf16(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), M, N, K,
                              _ptr(bias), _ptr(scale), _ptr(res), res.stride(0) if res is not None else 0, epi,
                              1 if out_f32 else 0, tile_cfg, _stream())
    _check(rc, "omc_gemm_bf16")
    return out


# ---- weight-streaming GEMM of the batched decode step (csrc/gemm_stream.cu): packed weights, stream-K, PDL, folded RMSNorm
PDL_ENABLED = os.environ.get("OMCHAT_B200_PDL", "1") != "0"
_stream_ws = {}


class PackedWeight:
    """An [N, K] weight re-laid as 16 KB swizzled tiles (omc_pack_weight); `data` is the 1024-byte aligned byte view."""

    def __init__(self, w: torch.Tensor, col_scale: Optional[torch.Tensor] = None):
        _need_cuda(w, col_scale)
        assert w.dim() == 2 and w.stride(1) == 1 and w.dtype == torch.bfloat16
        self.N, self.K = w.shape
        nbytes = load().omc_packed_weight_bytes(self.N, self.K)
        self._buf = torch.empty(nbytes + 1024, device=w.device, dtype=torch.uint8)
        off = (-self._buf.data_ptr()) % 1024
        self.data = self._buf[off:off + nbytes]
        rc = load().omc_pack_weight(_ptr(w), w.stride(0), self.N, self.K, _ptr(col_scale), self.data.data_ptr(), _stream())
        _check(rc, "omc_pack_weight")


def _stream_workspace(device) -> torch.Tensor:
    """Partial-tile slots + flags of the stream-K GEMM: consecutive launches of ONE stream share them (they are ordered),
    launches on different streams may overlap and get their own."""
    key = (str(device), torch.cuda.current_stream().cuda_stream)
    ws = _stream_ws.get(key)
    if ws is None:
        ws = torch.zeros(load().omc_gemm_stream_workspace_bytes(), device=device, dtype=torch.uint8)
        _stream_ws[key] = ws
    return ws


def ssq_parts(C: int) -> int:
    """Number of 128-column tiles of a C-wide row = number of sum-of-squares partials an EPI_RES epilogue writes."""
    return (C + 127) // 128


class TpXchg(ctypes.Structure):
    """Mirror of `omc_tp_xchg` (include/omchat_b200.h)."""
    _fields_ = [("rank", ctypes.c_int32), ("size", ctypes.c_int32), ("channel", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("bufs", c_void_p * 8)]


def tp_xchg(ptrs, rank: int, channel: int) -> TpXchg:
    """Exchange description of one row-parallel op: ptrs[r] = rank r's exchange buffer as mapped here (PeerExchange.ptrs)."""
    t = TpXchg()
    t.rank, t.size, t.channel = rank, len(ptrs), channel
    t.reserved = int(os.environ.get("OMCHAT_B200_XCHG_FENCE_ALL", "0"))  # A/B switch: per-thread system fence before the flag
    for r, p_ in enumerate(ptrs):
        t.bufs[r] = p_
    return t


def gemm_stream_xchg_bytes() -> int:
    return load().omc_gemm_stream_xchg_bytes()


def gemm_stream(x: torch.Tensor, wp: PackedWeight, out: Optional[torch.Tensor] = None, *, bias=None, res=None,
                epi: int = EPI_NONE, out_f32: bool = False, ssq_in: Optional[torch.Tensor] = None, ssq_in_parts: int = 0,
                norm_dim: int = 0, eps: float = 1e-6, ssq_out: Optional[torch.Tensor] = None, pdl: Optional[bool] = None,
                tp: Optional[TpXchg] = None):
    """out[M, N] = epi(rstd[m] * (x[M, K] @ W'^T)) for M <= 64 (see include/omchat_b200.h, omc_gemm_stream)."""
    _need_cuda(x, out, bias, res, ssq_in, ssq_out)
    M, K = x.shape
    assert K == wp.K and x.stride(1) == 1 and M <= 64
    N = wp.N
    n_out = N // 2 if epi == EPI_SWIGLU else N
    if out is None:
        out = torch.empty(M, n_out, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    assert out.shape == (M, n_out) and out.stride(1) == 1
    if ssq_in is not None:
        assert ssq_in.dtype == torch.float32 and ssq_in.numel() >= ssq_in_parts * 64 and norm_dim > 0
    if ssq_out is not None:
        assert ssq_out.dtype == torch.float32 and ssq_out.numel() >= ssq_parts(N) * 64
    ws = _stream_workspace(x.device)
    rc = load().omc_gemm_stream(_ptr(x), x.stride(0), M, wp.data.data_ptr(), N, K, _ptr(out), out.stride(0), 1 if out_f32 else 0,
                                _ptr(bias), _ptr(res), res.stride(0) if res is not None else 0, epi, _ptr(ssq_in),
                                ssq_in_parts, norm_dim, eps, _ptr(ssq_out), ws.data_ptr(),
                                1 if (PDL_ENABLED if pdl is None else pdl) else 0, ctypes.byref(tp) if tp is not None else None,
                                _stream())
    _check(rc, "omc_gemm_stream")
    return out


def row_ssq(x: torch.Tensor, ssq: torch.Tensor, parts: int = 1, pdl: Optional[bool] = None):
    _need_cuda(x, ssq)
    assert x.dim() == 2 and x.stride(1) == 1 and ssq.dtype == torch.float32 and ssq.numel() >= parts * 64
    rc = load().omc_row_ssq(_ptr(x), x.stride(0), x.shape[0], x.shape[1], _ptr(ssq), parts,
                            1 if (PDL_ENABLED if pdl is None else pdl) else 0, _stream())
    _check(rc, "omc_row_ssq")
    return ssq


def gemv(x: torch.Tensor, w: torch.Tensor, out: Optional[torch.Tensor] = None, *, norm_w=None, eps: float = 1e-6,
         bias=None, res=None, epi: int = EPI_NONE, out_f32: bool = False) -> torch.Tensor:
    _need_cuda(x, w)
    B, K = x.shape
    N = w.shape[0]
    n_out = N // 2 if epi == EPI_SWIGLU else N
    if out is None:
        out = torch.empty(B, n_out, device=x.device, dtype=torch.float32 if out_f32 else torch.bfloat16)
    rc = load().omc_gemv_bf16(_ptr(x), x.stride(0), _ptr(w), w.stride(0), _ptr(out), out.stride(0), B, N, K,
                              _ptr(norm_w), eps, _ptr(bias), _ptr(res), res.stride(0) if res is not None else 0, epi,
                              1 if out_f32 else 0, _stream())
    _check(rc, "omc_gemv_bf16")
    return out


def rmsnorm(x: torch.Tensor, w: torch.Tensor, eps: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(x, w)
    assert x.dim() == 2 and x.stride(1) == 1
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    rc = load().omc_rmsnorm(_ptr(x), x.stride(0), _ptr(w), _ptr(out), out.stride(0), x.shape[0], x.shape[1], eps,
                            _stream())
    _check(rc, "omc_rmsnorm")
    return out


def layernorm(x: torch.Tensor, w: torch.Tensor, b: Optional[torch.Tensor], eps: float,
              out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """torch.nn.LayerNorm over the last dim of bf16 rows (InternViT-300M norm_type = 'layer_norm')."""
    _need_cuda(x, w) if b is None else _need_cuda(x, w, b)
    assert x.dim() == 2 and x.stride(1) == 1
    if out is None:
        out = torch.empty(x.shape, device=x.device, dtype=torch.bfloat16)
    rc = load().omc_layernorm(_ptr(x), x.stride(0), _ptr(w), _ptr(b) if b is not None else None, _ptr(out), out.stride(0),
                              x.shape[0], x.shape[1], eps, _stream())
    _check(rc, "omc_layernorm")
    return out


def rmsnorm_pair(x: torch.Tensor, w_a: torch.Tensor, w_b: torch.Tensor, C: int, eps: float) -> torch.Tensor:
    """In place: x[:, :C] normalised with w_a, x[:, C:2C] with w_b (q_norm / k_norm on packed qkv rows), one launch."""
    _need_cuda(x, w_a, w_b)
    assert x.dim() == 2 and x.stride(1) == 1 and x.shape[1] >= 2 * C
    rc = load().omc_rmsnorm_pair(_ptr(x), x.stride(0), _ptr(w_a), _ptr(w_b), x.shape[0], C, eps, _stream())
    _check(rc, "omc_rmsnorm_pair")
    return x


def vit_im2col(pixels: torch.Tensor, ldc: int = 640) -> torch.Tensor:
    _need_cuda(pixels)
    assert pixels.dim() == 4 and pixels.shape[1] == 3 and pixels.is_contiguous()
    assert pixels.dtype in (torch.float32, torch.bfloat16)
    B, _, H, W = pixels.shape
    cols = torch.empty(B * (H // 14) * (W // 14), ldc, device=pixels.device, dtype=torch.bfloat16)
    rc = load().omc_vit_im2col(_ptr(pixels), 1 if pixels.dtype == torch.float32 else 0, _ptr(cols), ldc, B, H, W,
                               _stream())
    _check(rc, "omc_vit_im2col")
    return cols


def vit_assemble(patch: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, B: int) -> torch.Tensor:
    _need_cuda(patch, cls, pos)
    P = patch.shape[0] // B
    C = patch.shape[1]
    hidden = torch.empty(B * (P + 1), C, device=patch.device, dtype=torch.bfloat16)
    rc = load().omc_vit_assemble(_ptr(patch), _ptr(cls), _ptr(pos), _ptr(hidden), B, P, C, _stream())
    _check(rc, "omc_vit_assemble")
    return hidden


def select_pixel_shuffle(hidden: torch.Tensor, B: int, G: int, down: int) -> torch.Tensor:
    _need_cuda(hidden)
    C = hidden.shape[-1]
    out = torch.empty(B * (G // down) ** 2, C * down * down, device=hidden.device, dtype=torch.bfloat16)
    rc = load().omc_select_pixel_shuffle(_ptr(hidden), _ptr(out), B, G, C, down, _stream())
    _check(rc, "omc_select_pixel_shuffle")
    return out


def resample_u8(src: torch.Tensor, dst_h: int, dst_w: int, coefs: torch.Tensor, bounds: torch.Tensor, vertical: bool):
    """One separable Pillow-bicubic pass over an RGB uint8 image [H, W, 3] (omc_resample_u8)."""
    _need_cuda(src, coefs, bounds)
    assert src.dtype == torch.uint8 and src.dim() == 3 and src.shape[2] == 3 and src.is_contiguous()
    assert coefs.dtype == torch.int32 and bounds.dtype == torch.int32 and coefs.is_contiguous() and bounds.is_contiguous()
    dst = torch.empty(dst_h, dst_w, 3, device=src.device, dtype=torch.uint8)
    rc = load().omc_resample_u8(_ptr(src), src.shape[0], src.shape[1], _ptr(dst), dst_h, dst_w, _ptr(coefs), _ptr(bounds),
                                coefs.shape[1], 1 if vertical else 0, _stream())
    _check(rc, "omc_resample_u8")
    return dst


def anyres_pack(thumb, resized, target_w, target_h, paste_x, paste_y, crop, lut, dtype=torch.float32):
    _need_cuda(thumb, resized, lut)
    assert lut.dtype == torch.float32 and lut.shape == (3, 256) and lut.is_contiguous()
    n = 1 + (target_w // crop) * (target_h // crop)
    out = torch.empty(n, 3, crop, crop, device=thumb.device, dtype=dtype)
    rc = load().omc_anyres_pack(_ptr(thumb), _ptr(resized), resized.shape[1], resized.shape[0], target_w, target_h, paste_x,
                                paste_y, crop, _ptr(lut), _ptr(out), 1 if dtype == torch.bfloat16 else 0, _stream())
    _check(rc, "omc_anyres_pack")
    return out


def attention(q, k, v, out, cu_seqlens: torch.Tensor, max_seqlen: int, Hq: int, Hkv: int, causal: bool, scale: float,
              head_dim: int = 128):
    """q/k/v/out: 2-D row views [total, H*head_dim] (possibly slices of a packed qkv buffer: row stride = parent's);
    head_dim 128, or 64 (the InternViT-300M tower)."""
    _need_cuda(q, k, v, out, cu_seqlens)
    assert cu_seqlens.dtype == torch.int32
    rc = load().omc_attention_fwd_hd(_ptr(q), q.stride(0), _ptr(k), k.stride(0), _ptr(v), v.stride(0), _ptr(out),
                                     out.stride(0), _ptr(cu_seqlens), cu_seqlens.numel() - 1, max_seqlen,
                                     min(q.shape[0], k.shape[0], v.shape[0], out.shape[0]), Hq, Hkv, head_dim,
                                     1 if causal else 0, scale, _stream())
    _check(rc, "omc_attention_fwd")
    return out


def attention_set_impl(legacy):
    """False / 0 (default): tcgen05 kernel (64-key tiles, double-buffered S); True / 1: mma.sync kernel; 2: the first
    tcgen05 kernel (128-key tiles) — kept for A/B measurements."""
    rc = load().omc_attention_set_impl(int(legacy))
    if rc != 0:
        raise OmcError("omc_attention_set_impl failed")


def rope_kv_store(qkv, pos, seq_ids, Hq, Hkv, inv_freq, kv_pool, block_table, page_size):
    _need_cuda(qkv, pos, inv_freq, kv_pool, block_table)
    assert pos.dtype == torch.int32 and block_table.dtype == torch.int32 and inv_freq.dtype == torch.float32
    rc = load().omc_rope_kv_store(_ptr(qkv), qkv.stride(0), _ptr(pos), _ptr(seq_ids), qkv.shape[0], Hq, Hkv,
                                  _ptr(inv_freq), _ptr(kv_pool), _ptr(block_table), block_table.shape[1], page_size,
                                  _stream())
    _check(rc, "omc_rope_kv_store")


def decode_attn_splits(B: int, Hkv: int, max_ctx: int) -> int:
    return load().omc_decode_attn_splits(B, Hkv, max_ctx)


def decode_attn_workspace(B: int, Hq: int, Hkv: int, splits: int, device) -> torch.Tensor:
    n = load().omc_decode_attn_workspace_bytes(B, Hq, Hkv, splits)
    return torch.zeros((n + 3) // 4, device=device, dtype=torch.int32)


def paged_decode_attn(qkv, inv_freq, kv_pool, block_table, page_size, ctx_lens, Hq, Hkv, splits, scale, out, workspace):
    _need_cuda(qkv, kv_pool, block_table, ctx_lens, out, workspace)
    assert ctx_lens.dtype == torch.int32 and block_table.dtype == torch.int32
    rc = load().omc_paged_decode_attn(_ptr(qkv), qkv.stride(0), _ptr(inv_freq), _ptr(kv_pool), _ptr(block_table),
                                      block_table.shape[1], page_size, _ptr(ctx_lens), qkv.shape[0], Hq, Hkv, splits,
                                      scale, _ptr(out), out.stride(0), _ptr(workspace), _stream())
    _check(rc, "omc_paged_decode_attn")
    return out


def embed_lookup(ids: torch.Tensor, table: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    _need_cuda(ids, table)
    assert ids.dtype == torch.int64 and ids.is_contiguous()
    T, C = ids.numel(), table.shape[1]
    if out is None:
        out = torch.empty(T, C, device=table.device, dtype=torch.bfloat16)
    rc = load().omc_embed_lookup(_ptr(ids), T, _ptr(table), C, _ptr(out), out.stride(0), table.shape[0], _stream())
    _check(rc, "omc_embed_lookup")
    return out


def splice(ids: torch.Tensor, seq_offsets: torch.Tensor, table: torch.Tensor, feats: Optional[torch.Tensor],
           image_token: int, max_len: int, T_capacity: int):
    """Returns (embeds [T_capacity, C], pos_ids, seq_ids, out_offsets [n_seq+1]) — rows beyond out_offsets[-1] are
    unspecified."""
    _need_cuda(ids, seq_offsets, table)
    assert ids.dtype == torch.int64 and seq_offsets.dtype == torch.int32
    n_seq, S = seq_offsets.numel() - 1, ids.numel()
    C = table.shape[1]
    n_img, L = (feats.shape[0], feats.shape[1]) if feats is not None else (0, 1)
    dev = table.device
    embeds = torch.empty(T_capacity, C, device=dev, dtype=torch.bfloat16)
    pos_ids = torch.empty(T_capacity, device=dev, dtype=torch.int32)
    seq_ids = torch.empty(T_capacity, device=dev, dtype=torch.int32)
    out_off = torch.empty(n_seq + 1, device=dev, dtype=torch.int32)
    ws = torch.empty(S + n_seq + 2, device=dev, dtype=torch.int32)
    rc = load().omc_splice(_ptr(ids), _ptr(seq_offsets), n_seq, S, image_token, _ptr(table), _ptr(feats), n_img, L, C,
                           max_len, _ptr(embeds), _ptr(pos_ids), _ptr(seq_ids), _ptr(out_off), _ptr(ws), T_capacity,
                           table.shape[0], _stream())
    _check(rc, "omc_splice")
    return embeds, pos_ids, seq_ids, out_off


def argmax(logits: torch.Tensor, out: Optional[torch.Tensor] = None, workspace: Optional[torch.Tensor] = None):
    _need_cuda(logits)
    assert logits.dtype == torch.float32 and logits.dim() == 2 and logits.stride(1) == 1
    B, V = logits.shape
    if out is None:
        out = torch.empty(B, device=logits.device, dtype=torch.int64)
    if workspace is None:
        workspace = torch.empty(128 * B, device=logits.device, dtype=torch.float32)
    rc = load().omc_argmax(_ptr(logits), logits.stride(0), B, V, _ptr(out), _ptr(workspace), _stream())
    _check(rc, "omc_argmax")
    return out


# ----------------------------------------------------------------------------------------------- persistent decode step
_MEGA_PLAN_BYTES, _MEGA_OP_BYTES = 256, 96  # sizeof(MegaPlan), sizeof(MegaOp) in csrc/decode_mega.cu (profiling tools only)


def mega_op_kinds(plan) -> list:
    """Names of the ops of a built decode plan, for tools/prof_mega.py: parsed from the host copy of the op list
    (MegaOp: 7 pointers, N, K, ldx, ldo, kc0, ldw, int16 in_op, then uint8 type, epi, R, ksplit, gran, flags, ...)."""
    raw, names, prev_attn = plan.host.raw, [], False
    for i in range(plan.n_ops):
        o = _MEGA_PLAN_BYTES + i * _MEGA_OP_BYTES
        typ, epi, flags = raw[o + 82], raw[o + 83], raw[o + 87]
        if typ == 2:
            name = "attn"
        elif typ == 3:
            name = "final"
        elif flags & 1:
            name = "lm_head"
        elif epi == 3:
            name = "gate_up"
        elif epi in (8, 9) or (epi == 2 and not prev_attn):
            name = "down"
        elif epi == 2:
            name = "o"
        else:
            name = "qkv"
        prev_attn = typ == 2
        names.append(name)
    return names


class DecodeDesc(ctypes.Structure):
    """Mirror of `omc_decode_desc` (include/omchat_b200.h)."""
    _fields_ = ([(n, ctypes.c_int32) for n in ("n_layers", "batch", "hidden", "q_heads", "kv_heads", "inter", "vocab",
                                                 "vocab_offset", "page_size", "max_pages", "grid", "hist_capacity",
                                                 "rope_positions", "l2_prefetch_stages", "ring_slot_bytes", "tune")]
                + [("eps", c_float), ("attn_scale", c_float)]
                + [(n, c_void_p) for n in ("embed", "final_norm", "lm_head", "rope_cs", "ln1", "qkv_w", "qkv_b", "o_w",
                                           "ln2", "gate_up_w", "down_w", "kv_pool")]
                + [("kv_layer_stride", c_longlong)]
                + [(n, c_void_p) for n in ("block_table", "ctx_lens", "tokens", "token_hist", "hist_pos", "h", "qkv",
                                           "attn", "act", "logits", "workspace", "status", "prof")]
                + [("tp_rank", ctypes.c_int32), ("tp_size", ctypes.c_int32), ("xchg", c_void_p * 8)])


def num_sms() -> int:
    n = load().omc_num_sms()
    if n <= 0:
        raise OmcError("no CUDA device (omc_num_sms)")
    return n


class PeerExchange:
    """One exchange buffer per tensor-parallel rank, each mapped into every rank's address space (CUDA IPC over NVLink):
    the persistent decode kernel's replacement for the NCCL communicator (include/omchat_b200.h, omc_peer_*).
    `group` is the torch.distributed group whose ranks share the node; the 64-byte handles travel through it."""

    def __init__(self, nbytes: int, rank: int, size: int, group=None):
        import torch.distributed as dist
        lib = load()
        self.rank, self.size, self.nbytes = rank, size, nbytes
        own = c_void_p()
        handle = ctypes.create_string_buffer(64)
        rc = lib.omc_peer_alloc(nbytes, ctypes.byref(own), handle)
        if rc != 0:
            raise OmcError(f"omc_peer_alloc failed ({rc}): {lib.omc_last_error().decode(errors='replace')}")
        self.own = own.value
        handles = [None] * size
        dist.all_gather_object(handles, bytes(handle.raw), group=group)
        self.ptrs = []
        for p in range(size):
            if p == rank:
                self.ptrs.append(self.own)
                continue
            ptr = c_void_p()
            rc = lib.omc_peer_open(handles[p], ctypes.byref(ptr))
            if rc != 0:
                raise OmcError(f"omc_peer_open(rank {p}) failed ({rc}): {lib.omc_last_error().decode(errors='replace')}")
            self.ptrs.append(ptr.value)
        dist.barrier(group=group)

    def close(self):
        lib = load()
        for p, ptr in enumerate(self.ptrs):
            if p != self.rank and ptr:
                lib.omc_peer_close(ptr)
        if self.own:
            lib.omc_peer_free(self.own)
        self.ptrs, self.own = [], None


def decode_xchg_bytes(batch: int, hidden: int, tp_size: int) -> int:
    d = DecodeDesc()
    d.batch, d.hidden, d.tp_size = batch, hidden, tp_size
    n = load().omc_decode_xchg_bytes(ctypes.byref(d))
    if n <= 0:
        raise OmcError("omc_decode_xchg_bytes failed")
    return n


class DecodePlan:
    """Host + device copies of one megakernel plan (omc_decode_plan_build) and the buffers it points at. The plan bakes in
    every pointer (weights, KV pool, block table, state), so it must be rebuilt when any of them is reallocated."""

    def __init__(self, *, layers, embed, final_norm, lm_head, rope_cs, cfg_dims, kv_pool, block_table, ctx_lens, tokens,
                 token_hist, hist_pos, h, qkv, attn, act, logits, page_size, eps, scale, vocab_offset=0, grid=None,
                 tp_rank=0, tp_size=1, xchg_ptrs=None):
        lib = load()
        n_layers = len(layers)
        B = tokens.numel()
        dev = tokens.device
        self.grid = grid or num_sms()
        self.epoch = 1
        arr = lambda ts: (c_void_p * max(n_layers, 1))(*[t.data_ptr() for t in ts])  # noqa: E731
        self._arrays = [arr([getattr(l, k) for l in layers]) for k in
                        ("ln1", "qkv_w", "qkv_b", "o_w", "ln2", "gate_up_w", "down_w")]
        for l in layers:
            for k in ("qkv_w", "o_w", "gate_up_w", "down_w"):
                assert getattr(l, k).is_contiguous()
        assert lm_head.is_contiguous() and kv_pool.is_contiguous()
        d = DecodeDesc()
        d.n_layers, d.batch = n_layers, B
        d.hidden, d.q_heads, d.kv_heads, d.inter, d.vocab = cfg_dims
        d.vocab_offset, d.page_size, d.max_pages = vocab_offset, page_size, block_table.shape[1]
        d.grid, d.hist_capacity = self.grid, (token_hist.shape[0] if token_hist is not None else 0)
        d.eps, d.attn_scale = eps, scale
        d.l2_prefetch_stages = int(os.environ.get("OMCHAT_B200_MEGA_PF", "0"))
        d.ring_slot_bytes = int(os.environ.get("OMCHAT_B200_MEGA_SLOT", "0"))  # 0 = the library's default
        d.tune = int(os.environ.get("OMCHAT_B200_MEGA_TUNE", "0"))  # A/B switches of omc_decode_desc.tune (measurements)
        d.embed, d.final_norm, d.lm_head, d.rope_cs = embed.data_ptr(), final_norm.data_ptr(), lm_head.data_ptr(), rope_cs.data_ptr()
        assert rope_cs.dtype == torch.float32 and rope_cs.is_contiguous() and rope_cs.shape[1:] == (64, 2)
        d.rope_positions = rope_cs.shape[0]
        (d.ln1, d.qkv_w, d.qkv_b, d.o_w, d.ln2, d.gate_up_w, d.down_w) = [ctypes.cast(a, c_void_p) for a in self._arrays]
        d.kv_pool = kv_pool.data_ptr()
        d.kv_layer_stride = kv_pool.stride(0) if n_layers > 0 else 0
        d.block_table, d.ctx_lens, d.tokens = block_table.data_ptr(), ctx_lens.data_ptr(), tokens.data_ptr()
        d.token_hist = token_hist.data_ptr() if token_hist is not None else None
        d.hist_pos = hist_pos.data_ptr() if hist_pos is not None else None
        d.h, d.qkv, d.attn, d.act, d.logits = h.data_ptr(), qkv.data_ptr(), attn.data_ptr(), act.data_ptr(), logits.data_ptr()
        d.tp_rank, d.tp_size = tp_rank, tp_size
        if tp_size > 1:
            assert xchg_ptrs is not None and len(xchg_ptrs) == tp_size
            for p, ptr in enumerate(xchg_ptrs):
                d.xchg[p] = ptr
        ws_bytes = lib.omc_decode_workspace_bytes(ctypes.byref(d))
        if ws_bytes <= 0:
            raise OmcError("omc_decode_workspace_bytes failed")
        self.workspace = torch.zeros(ws_bytes, device=dev, dtype=torch.uint8)
        d.workspace = self.workspace.data_ptr()
        self.status = torch.zeros(4, dtype=torch.int32).pin_memory()  # watchdog record, readable after a device trap
        d.status = self.status.data_ptr()
        nbytes = lib.omc_decode_plan_bytes(n_layers)
        self.prof = None
        if os.environ.get("OMCHAT_B200_MEGA_PROF", "0") == "1":
            # sized for the longest op list the plan builder may produce; viewed as [grid, n_ops, 8] once n_ops is known
            self._prof_flat = torch.zeros(self.grid * (nbytes // _MEGA_OP_BYTES + 1) * 8, device=dev, dtype=torch.int64)
            d.prof = self._prof_flat.data_ptr()
        self.desc = d
        self.host = ctypes.create_string_buffer(nbytes)
        rc = lib.omc_decode_plan_build(ctypes.byref(d), self.host)
        if rc != 0:
            raise OmcError(f"omc_decode_plan_build failed ({rc}): {lib.omc_last_error().decode(errors='replace')}")
        self.dev = torch.frombuffer(bytearray(self.host.raw), dtype=torch.uint8).to(dev)
        self.n_ops = int.from_bytes(self.host.raw[0:4], "little")  # MegaPlan.n_ops
        if d.prof:
            self.prof = self._prof_flat[: self.grid * self.n_ops * 8].view(self.grid, self.n_ops, 8)
        self._keep = (layers, embed, final_norm, lm_head, rope_cs, kv_pool, block_table, ctx_lens, tokens, token_hist,
                      hist_pos, h, qkv, attn, act, logits)

    def step(self, epoch: Optional[int] = None):
        """One decode step. `epoch` tags the in-flight activations: by default a per-plan counter; under tensor parallelism
        the caller passes a counter shared by every plan that uses the same peer exchange buffers."""
        if epoch is None:
            epoch = self.epoch
            self.epoch += 1
        rc = load().omc_decode_step(self.host, self.dev.data_ptr(), epoch & 0xFFFFFF, _stream())
        _check(rc, "omc_decode_step")


# ----------------------------------------------------------------------------------------------- Qwen2-MoE sparse block
class MoeWorkspace:
    """Device buffers of one sparse-MoE block call for up to T tokens (sized for the worst-case routing, reused by every
    layer): routing results, the expert-sorted copies of the rows, the experts' activations and outputs."""

    def __init__(self, T: int, C: int, n_experts: int, top_k: int, moe_inter: int, shared_inter: int, device):
        self.T, self.C, self.E, self.k = T, C, n_experts, top_k
        self.max_tiles = load().omc_moe_max_tiles(T, top_k, n_experts)
        M = self.max_tiles * 128
        i32 = dict(device=device, dtype=torch.int32)
        self.topk_ids = torch.zeros(T, top_k, **i32)
        self.topk_w = torch.zeros(T, top_k, device=device, dtype=torch.float32)
        self.shared_gate = torch.zeros(T, device=device, dtype=torch.float32)
        self.counts = torch.zeros(n_experts, **i32)  # zero on entry of every route call (omc_moe_plan re-zeroes it)
        self.seg_start = torch.zeros(n_experts, **i32)
        self.cursor = torch.zeros(n_experts, **i32)
        self.tile_expert = torch.full((self.max_tiles,), -1, **i32)
        self.slot_of = torch.zeros(T, top_k, **i32)
        bf = dict(device=device, dtype=torch.bfloat16)
        self.xperm = torch.zeros(M, C, **bf)  # zero once: padding rows are multiplied (never read back) - keep them finite
        self.aperm = torch.zeros(M, moe_inter, **bf)
        self.yperm = torch.zeros(M, C, **bf)
        self.logits = None  # fp32 [T, 128 n] router logits of the tensor-core router path (allocated on first use)
        self.shared_act = torch.empty(T, shared_inter, **bf) if shared_inter > 0 else None
        self.shared_y = torch.empty(T, C, **bf) if shared_inter > 0 else None


ROUTER_GEMM_MIN_T = 256  # from this many tokens on the router logits are computed on the tensor cores


def router_cat(router_w: torch.Tensor, shared_gate_w: Optional[torch.Tensor]) -> torch.Tensor:
    """[router_w; shared_gate_w; zero rows] padded to a multiple of 128 rows: the B operand of the router-logits GEMM."""
    E, C = router_w.shape
    rows = E + (1 if shared_gate_w is not None else 0)
    out = torch.zeros((rows + 127) // 128 * 128, C, device=router_w.device, dtype=torch.bfloat16)
    out[:E] = router_w
    if shared_gate_w is not None:
        out[E] = shared_gate_w.view(-1)
    return out


def moe_block(h: torch.Tensor, xn: torch.Tensor, ws: MoeWorkspace, router_w: torch.Tensor, shared_gate_w: Optional[torch.Tensor],
              experts_gate_up: torch.Tensor, experts_down: torch.Tensor, shared_gate_up: Optional[torch.Tensor],
              shared_down: Optional[torch.Tensor], norm_topk: bool, norm_w: Optional[torch.Tensor] = None,
              eps: float = 1e-6, shared_y: Optional[torch.Tensor] = None, ssq_out: Optional[torch.Tensor] = None,
              ssq_parts: int = 1, router_cat_w: Optional[torch.Tensor] = None, defer_combine: bool = False) -> torch.Tensor:
    """h[T, C] += SparseMoeBlock(xn[T, C]) (transformers modeling_qwen2_moe.py:363-374), in place. experts_gate_up
    [E * 2 I, C] with every expert's gate / up rows interleaved (the SwiGLU epilogue's layout), experts_down [E * C, I].
    norm_w given: xn is an OUTPUT - the router kernel computes xn = RMSNorm(h) * norm_w itself (one launch less).
    shared_y given: the shared expert's output [T, C], already computed by the caller (the decode step runs it on the
    weight-streaming GEMMs); shared_gate_up / shared_down are then unused.
    router_cat_w given (see router_cat()) and T >= ROUTER_GEMM_MIN_T: the E + 1 logits per token are computed on the tensor
    cores - RMSNorm kernel -> omc_gemm_bf16 with fp32 output -> omc_moe_select - instead of on the CUDA cores.
    defer_combine: stop after the routed experts' GEMMs; the caller finishes with moe_combine()."""
    _need_cuda(h, xn, router_w, experts_gate_up, experts_down)
    T, C = xn.shape
    assert T <= ws.T and C == ws.C and h.shape == xn.shape and xn.stride(1) == 1 and h.stride(1) == 1
    E, k, L, st = ws.E, ws.k, load(), _stream()
    I2 = experts_gate_up.shape[0] // E
    max_tiles = L.omc_moe_max_tiles(T, k, E)  # tiles this call can touch (<= the workspace's)
    if router_cat_w is not None and T >= ROUTER_GEMM_MIN_T:
        if norm_w is not None:
            rmsnorm(h, norm_w, eps, out=xn)
        W = router_cat_w.shape[0]
        if ws.logits is None or ws.logits.shape[1] != W:
            ws.logits = torch.empty(ws.T, W, device=xn.device, dtype=torch.float32)
        gemm(xn, router_cat_w, out=ws.logits[:T], out_f32=True)
        _check(L.omc_moe_select(_ptr(ws.logits), W, T, E, k, int(norm_topk), 1 if shared_gate_w is not None else 0,
                                _ptr(ws.topk_ids), _ptr(ws.topk_w), _ptr(ws.shared_gate), _ptr(ws.counts), st), "omc_moe_select")
    else:
        src = h if norm_w is not None else xn
        _check(L.omc_moe_route(_ptr(src), src.stride(0), T, C, _ptr(norm_w), eps, _ptr(xn) if norm_w is not None else None,
                               xn.stride(0), _ptr(router_w), _ptr(shared_gate_w), E, k, int(norm_topk), _ptr(ws.topk_ids),
                               _ptr(ws.topk_w), _ptr(ws.shared_gate), _ptr(ws.counts), st), "omc_moe_route")
    _check(L.omc_moe_plan_scatter(_ptr(ws.counts), E, ws.max_tiles, _ptr(ws.seg_start), _ptr(ws.cursor), _ptr(ws.tile_expert),
                                  _ptr(xn), xn.stride(0), T, C, _ptr(ws.topk_ids), k, _ptr(ws.xperm), C, _ptr(ws.slot_of), st),
           "omc_moe_plan_scatter")
    Mc = max_tiles * 128
    hint = min(max_tiles, T * k)
    _check(L.omc_gemm_bf16_grouped(_ptr(ws.xperm), C, Mc, _ptr(experts_gate_up), C, E, I2, C, _ptr(ws.tile_expert), hint,
                                   _ptr(ws.aperm), I2 // 2, EPI_SWIGLU, st), "omc_gemm_bf16_grouped")
    _check(L.omc_gemm_bf16_grouped(_ptr(ws.aperm), I2 // 2, Mc, _ptr(experts_down), I2 // 2, E, C, I2 // 2,
                                   _ptr(ws.tile_expert), hint, _ptr(ws.yperm), C, EPI_NONE, st), "omc_gemm_bf16_grouped")
    add_launches(3)
    if defer_combine:
        return h
    if shared_y is None and shared_gate_up is not None:
        gemm(xn, shared_gate_up, out=ws.shared_act[:T], epi=EPI_SWIGLU)
        shared_y = gemm(ws.shared_act[:T], shared_down, out=ws.shared_y[:T])
    return moe_combine(h, ws, shared_y, ssq_out, ssq_parts)


def moe_combine(h: torch.Tensor, ws: MoeWorkspace, shared_y: Optional[torch.Tensor], ssq_out: Optional[torch.Tensor] = None,
                ssq_parts: int = 1) -> torch.Tensor:
    """The last step of moe_block (for callers that ran it with defer_combine=True and computed the shared expert on another
    stream meanwhile): h += sum_j w_j * y[slot_j] + sigmoid_gate * shared_y."""
    T, C = h.shape
    k = ws.k
    _check(load().omc_moe_combine(_ptr(h), h.stride(0), T, C, _ptr(ws.yperm), C, _ptr(ws.slot_of), _ptr(ws.topk_w), k,
                                  _ptr(shared_y), shared_y.stride(0) if shared_y is not None else C, _ptr(ws.shared_gate),
                                  _ptr(ssq_out), ssq_parts, _stream()), "omc_moe_combine")
    add_launches(1 if T * k <= 16 else 2)
    return h


# ----------------------------------------------------------------------------------------------- model-level entry points
class VitDesc(ctypes.Structure):
    """Mirror of `omc_vit_desc` (include/omchat_b200.h)."""
    _fields_ = ([(n, ctypes.c_int32) for n in ("n_layers", "hidden", "heads", "inter", "image_size", "patch_size", "patch_k",
                                                 "qk_norm", "pixel_shuffle_down", "proj_hidden")]
                + [("eps", c_float), ("norm_folded", ctypes.c_int32)]
                + [(n, c_void_p) for n in ("patch_w", "patch_b", "cls", "pos", "norm1", "qkv_w", "q_norm", "k_norm", "proj_w",
                                           "proj_b", "ls1", "norm2", "fc1_w", "fc1_b", "fc2_w", "fc2_b", "ls2", "p_w0", "p_b0",
                                           "p_w2", "p_b2")]
                + [("norm_type", ctypes.c_int32), ("attn_head_dim", ctypes.c_int32), ("norm1_b", c_void_p), ("norm2_b", c_void_p),
                   ("qkv_b", c_void_p)])


class VitForward:
    """omc_vit_forward on a model's weights: the whole encode_images (tower + select / pixel shuffle + projector) as ONE
    C call - what a non-Python host binds. `vit` / `proj` are weights.VitW / ProjW, vc an InternVisionConfig."""

    def __init__(self, vit, proj, vc, pixel_shuffle_down: int = 1, folded=None, mats=None, mats_folded: bool = False):
        """folded: [(qkv * norm1, fc1 * norm2)] per layer (InternVITVisionTower._folded()) -> the norm-folded loop the
        product runs by default; None -> plain weights + stand-alone norm kernels. mats: the tower's _layer_mats() - per layer
        (qkv_w, qkv_b, proj_w, fc1_w) with heads zero-padded to 128 dims - required when head_dim < 128 (InternViT-300M)."""
        n = len(vit.layers)
        if mats is not None:
            folded = None
        elif vc.head_dim not in (64, 128):
            raise ValueError("head_dim other than 64 / 128: pass mats=tower._layer_mats() (zero-padded heads)")
        padded = mats is not None and mats[0][2].shape[1] != vc.hidden_size  # proj_w widened: heads zero-padded to 128
        d = VitDesc()
        d.norm_folded = 1 if (folded is not None or (mats is not None and mats_folded)) else 0
        self._folded = folded
        d.n_layers, d.hidden, d.heads, d.inter = n, vc.hidden_size, vc.num_attention_heads, vc.intermediate_size
        d.image_size, d.patch_size, d.patch_k, d.qk_norm = vc.image_size, vc.patch_size, vit.patch_w.shape[1], int(vc.qk_normalization)
        d.pixel_shuffle_down, d.proj_hidden, d.eps = pixel_shuffle_down, proj.w2.shape[0], vc.layer_norm_eps
        d.patch_w, d.patch_b, d.cls, d.pos = (t.data_ptr() for t in (vit.patch_w, vit.patch_b, vit.cls, vit.pos))
        names = {"norm1": "norm1", "qkv_w": "qkv", "q_norm": "q_norm", "k_norm": "k_norm", "proj_w": "proj_w", "proj_b": "proj_b",
                 "ls1": "ls1", "norm2": "norm2", "fc1_w": "fc1_w", "fc1_b": "fc1_b", "fc2_w": "fc2_w", "fc2_b": "fc2_b", "ls2": "ls2"}
        self._arrays = {}
        self._mats = mats
        extra = {"norm1_b": "norm1_b", "norm2_b": "norm2_b"} if vc.norm_type == "layer_norm" else {}
        if any(l.qkv_b is not None for l in vit.layers):
            extra["qkv_b"] = "qkv_b"
        d.norm_type = 1 if vc.norm_type == "layer_norm" else 0
        d.attn_head_dim = 128 if padded else 0
        for field, attr in {**names, **extra}.items():
            if field in ("q_norm", "k_norm") and not vc.qk_normalization:
                continue
            if folded is not None and field in ("qkv_w", "fc1_w"):
                ptrs = [f[0 if field == "qkv_w" else 1].data_ptr() for f in folded]
            elif mats is not None and field in ("qkv_w", "qkv_b", "proj_w", "fc1_w"):
                ptrs = [m[("qkv_w", "qkv_b", "proj_w", "fc1_w").index(field)].data_ptr() for m in mats]
            else:
                ptrs = [getattr(l, attr).data_ptr() for l in vit.layers]
            arr = (c_void_p * max(n, 1))(*ptrs)
            self._arrays[field] = arr
            setattr(d, field, ctypes.cast(arr, c_void_p))
        d.p_w0, d.p_b0, d.p_w2, d.p_b2 = (t.data_ptr() for t in (proj.w0, proj.b0, proj.w2, proj.b2))
        self.desc, self._keep, self._ws = d, (vit, proj), None
        self.tokens = (vc.image_size // vc.patch_size // pixel_shuffle_down) ** 2

    def __call__(self, pixels: torch.Tensor) -> torch.Tensor:
        _need_cuda(pixels)
        assert pixels.dim() == 4 and pixels.is_contiguous() and pixels.dtype in (torch.float32, torch.bfloat16)
        n = pixels.shape[0]
        need = load().omc_vit_workspace_bytes(ctypes.byref(self.desc), n)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty(need, device=pixels.device, dtype=torch.uint8)
        out = torch.empty(n, self.tokens, self.desc.proj_hidden, device=pixels.device, dtype=torch.bfloat16)
        rc = load().omc_vit_forward(ctypes.byref(self.desc), _ptr(pixels), 1 if pixels.dtype == torch.float32 else 0, n,
                                    self._ws.data_ptr(), _ptr(out), _stream())
        per_layer = 7 + (2 if self.desc.qk_norm else 0) - (2 + (1 if self.desc.qk_norm else 0)) * self.desc.norm_folded
        add_launches(4 + self.desc.norm_folded + self.desc.n_layers * per_layer + 3 - 1)
        _check(rc, "omc_vit_forward")
        return out


def decoder_prefill(layers, final_norm, lm_head, dims, eps, scale, inv_freq, embeds, pos_ids, seq_ids, cu_seqlens, max_len,
                    kv_pool, block_table, page_size, last_rows=None, folded=None):
    """omc_decoder_prefill: embeds [T, C] (updated in place) -> fp32 logits [n_seq, V] of each sequence's last row (None if
    last_rows is None); fills the paged cache. dims = (hidden, q_heads, kv_heads, inter, vocab)."""
    _need_cuda(embeds, pos_ids, seq_ids, cu_seqlens, kv_pool, block_table, inv_freq)
    n_layers, T, n_seq = len(layers), embeds.shape[0], cu_seqlens.numel() - 1
    d = DecodeDesc()
    d.n_layers = n_layers
    d.hidden, d.q_heads, d.kv_heads, d.inter, d.vocab = dims
    d.page_size, d.max_pages, d.eps, d.attn_scale = page_size, block_table.shape[1], eps, scale
    d.final_norm, d.lm_head = final_norm.data_ptr(), lm_head.data_ptr()
    def ptrs(k):
        if folded is not None and k in ("qkv_w", "gate_up_w"):  # [(q|k|v * ln1, gate|up * ln2)] per layer
            return [f[0 if k == "qkv_w" else 1].data_ptr() for f in folded]
        return [getattr(l, k).data_ptr() for l in layers]

    arrays = [(c_void_p * max(n_layers, 1))(*ptrs(k)) for k in ("ln1", "qkv_w", "qkv_b", "o_w", "ln2", "gate_up_w", "down_w")]
    (d.ln1, d.qkv_w, d.qkv_b, d.o_w, d.ln2, d.gate_up_w, d.down_w) = [ctypes.cast(a, c_void_p) for a in arrays]
    d.kv_pool, d.kv_layer_stride, d.block_table = kv_pool.data_ptr(), kv_pool.stride(0), block_table.data_ptr()
    ws = torch.empty(load().omc_decoder_prefill_workspace_bytes(ctypes.byref(d), T, n_seq), device=embeds.device, dtype=torch.uint8)
    logits = torch.empty(n_seq, dims[4], device=embeds.device, dtype=torch.float32) if last_rows is not None else None
    rc = load().omc_decoder_prefill(ctypes.byref(d), _ptr(inv_freq), _ptr(embeds), _ptr(pos_ids), _ptr(seq_ids), _ptr(cu_seqlens),
                                    n_seq, T, max_len, _ptr(last_rows), ws.data_ptr(), _ptr(logits), 1 if folded is not None else 0,
                                    _stream())
    add_launches((6 if folded is not None else 8) * n_layers + (1 if folded is not None else 0) + (3 if last_rows is not None else 0) - 1)
    _check(rc, "omc_decoder_prefill")
    return logits
