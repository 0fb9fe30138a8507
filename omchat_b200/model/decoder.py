"""Qwen2-7B decoder (prefill + greedy decode over a paged KV cache) on the C-ABI kernels.

The decoder arithmetic is not in the reference tree: OmChatQwen2ForCausalLM subclasses transformers' Qwen2ForCausalLM
(omchat/model/language_model/omchat_qwen2.py:7,29,77; pin transformers==4.41.2, pyproject.toml:22). What is mirrored
here is transformers models/qwen2/modeling_qwen2.py: Qwen2DecoderLayer.forward :280-310, Qwen2Attention.forward
:206-246 (q/k/v with bias, rotate-half RoPE :124-146, GQA :149-158, cache update :227), Qwen2MLP :46-48, final norm
:411 and lm_head :470-472.

Data layout in HBM
  * tokens of all sequences of a prefill are PACKED: activations are [T_total, C] bf16, sequence s owns rows
    offsets[s]..offsets[s+1]; attention runs var-len over cu_seqlens, so padding never reaches a kernel
  * q|k|v of a layer share one [T, (Hq+2*Hkv)*128] buffer (one GEMM, fused bias); gate|up share one weight matrix with
    rows alternating gate_i, up_i so SwiGLU is a GEMM/GEMV epilogue
  * KV cache: one pool per layer [num_pages, 2, Hkv, page_size, 128] bf16 + an int32 block table [n_seq, max_pages]
    (replaces DynamicCache / torch.cat per step, modeling_qwen2.py:227)
Per layer, prefill = rmsnorm, qkv GEMM(+bias), RoPE+KV-append, causal flash attention, o GEMM(+residual), rmsnorm,
gate/up GEMM(SwiGLU), down GEMM(+residual). Decode (B <= 8) = 5 launches: qkv GEMV (RMSNorm fused), paged attention
(RoPE + append fused), o GEMV(+residual), gate/up GEMV (RMSNorm + SwiGLU fused), down GEMV(+residual); the whole step
including lm_head, argmax and the next embedding lookup is captured in one CUDA graph.

Batched decode (B > 4) = per layer 5 launches chained by programmatic dependent launch (csrc/gemm_stream.cu): qkv GEMM
(RMSNorm folded in: packed weights carry the norm weight, the epilogue applies rstd), paged attention (RoPE + append
fused), o GEMM (+residual, emits the row sums of squares), gate/up GEMM (folded RMSNorm + SwiGLU), down GEMM (+residual,
sums of squares); one CUDA graph per step.

Tensor parallelism (tp > 1, one process per GPU): q/k/v and gate/up are column-parallel, o and down row-parallel with
one NCCL all-reduce each (56 per forward); rank 0 alone folds the residual into its partial sum so the all-reduce
result is the new residual stream. lm_head is vocab-parallel; greedy sampling all-gathers one (max, index) pair per rank.
"""
from __future__ import annotations

import math
import os
from dataclasses import dataclass
from typing import List, Optional, Sequence

import torch

from .. import lib
from ..config import OmChatQwen2Config
from .weights import LlmW

GEMV_MAX_B = 8
GEMV_MAX_SMEM = 200 * 1024
MEGA_MAX_B = 4  # the persistent decode kernel (csrc/decode_mega.cu) handles 1..4 sequences
# Smallest batch that decodes on the weight-streaming GEMMs (csrc/gemm_stream.cu); below it: the persistent kernel
# (csrc/decode_mega.cu). Measured on one B200, ctx 1024 (profiles/r02_decode_small_batch_stream_vs_mega.jsonl), ms per step,
# persistent kernel / streaming GEMMs: batch 1: 2.68 / 2.82, batch 2: 3.18 / 2.85, batch 4: 5.24 / 2.89 - the persistent
# kernel re-reads its activation fragments for every weight byte from batch 2 on, the tcgen05 tile does not care.
# Under tensor parallelism the persistent kernel keeps batches 1..4 (its in-kernel exchange moves 8-byte words, not tiles).
STREAM_MIN_B = 2
STREAM_MIN_B_TP = 5
MEGA_HIST = 4096  # token-history rows kept on the device between host reads


def rope_inv_freq(cfg: OmChatQwen2Config, device) -> torch.Tensor:
    """Qwen2RotaryEmbedding default init (modeling_qwen2.py:51-100): computed on the host in fp32 exactly as HF does."""
    d = cfg.head_dim
    inv = 1.0 / (cfg.rope_theta ** (torch.arange(0, d, 2, dtype=torch.int64).to(torch.float32) / d))
    return inv.to(device)


class PagedKVCache:
    """Paged replacement for the reference's DynamicCache / tuple cache (omchat_arch.py:63, modeling_qwen2.py:227).

    pool[layer] is [num_pages, 2, Hkv, page_size, 128]; sequence s stores position p in page
    block_table[s, p // page_size], slot p % page_size. `shuffle_pages` permutes the page ids so that block tables are
    non-contiguous (exercises paging)."""

    def __init__(self, n_layers: int, n_seq: int, max_ctx: int, kv_heads: int, page_size: int, device,
                 shuffle_pages: bool = True, seed: int = 0):
        self.page_size = page_size
        self.n_seq = n_seq
        self.max_pages = (max_ctx + page_size - 1) // page_size
        self.capacity = self.max_pages * page_size
        n_pages = n_seq * self.max_pages
        self.pool = torch.zeros(n_layers, n_pages, 2, kv_heads, page_size, 128, device=device, dtype=torch.bfloat16)
        if shuffle_pages:
            g = torch.Generator().manual_seed(seed)
            ids = torch.randperm(n_pages, generator=g)
        else:
            ids = torch.arange(n_pages)
        self.block_table = ids.to(torch.int32).view(n_seq, self.max_pages).to(device)
        self.ctx_lens = torch.zeros(n_seq, device=device, dtype=torch.int32)  # device copy used by the kernels
        self.host_lens = [0] * n_seq

    # HF-cache-like helpers used by callers of the reference API
    def get_seq_length(self, layer_idx: int = 0) -> int:
        return max(self.host_lens) if self.host_lens else 0

    def __len__(self):
        return self.pool.shape[0]

    def __bool__(self):  # `if past_key_values:` in prepare_inputs_for_generation (omchat_qwen2.py:95)
        return self.get_seq_length() > 0

    def gather(self, layer: int, seq: int):
        """Contiguous (K, V) [Hkv, ctx, 128] of one sequence — for tests / debugging only."""
        n = self.host_lens[seq]
        pages = self.block_table[seq, : (n + self.page_size - 1) // self.page_size].long()
        kv = self.pool[layer, pages]  # [p, 2, Hkv, page, 128]
        kv = kv.permute(1, 2, 0, 3, 4).reshape(2, kv.shape[2], -1, 128)[:, :, :n]
        return kv[0], kv[1]


@dataclass
class TPInfo:
    rank: int = 0
    size: int = 1
    group: object = None


class Qwen2Decoder:
    def __init__(self, cfg: OmChatQwen2Config, w: LlmW, tp: Optional[TPInfo] = None):
        self.cfg = cfg
        self.w = w
        self.tp = tp or TPInfo()
        self.device = w.norm.device
        self.inv_freq = rope_inv_freq(cfg, self.device)
        self.C = cfg.hidden_size
        # local (per TP rank) head counts are read off the sharded weights
        self.Hkv = w.kv_heads_local
        self.Hq = w.q_heads_local
        dense = [l for l in w.layers if l.gate_up_w is not None]
        self.I_local = dense[0].gate_up_w.shape[0] // 2 if dense else 8  # 8: placeholder width when every layer is sparse
        self.V_local = w.lm_head.shape[0]
        self.scale = cfg.head_dim ** -0.5
        self.eps = cfg.rms_norm_eps
        self.mega_enabled = os.environ.get("OMCHAT_B200_NO_MEGA", "0") != "1"
        # batched decode steps (B > 4): weight-streaming GEMMs on packed weights (0 = the round-1 skinny GEMM / GEMV path)
        self.stream_enabled = os.environ.get("OMCHAT_B200_NO_STREAM", "0") != "1"
        self.stream_min_b = int(os.environ.get("OMCHAT_B200_STREAM_MIN_B",
                                               str(STREAM_MIN_B if self.tp.size == 1 else STREAM_MIN_B_TP)))
        # prefill: RMSNorms folded into the GEMMs that follow them (0 = stand-alone RMSNorm kernels, the round-1 path)
        self.fold_norms = os.environ.get("OMCHAT_B200_FOLD_NORMS", "1") != "0"
        self._folded_w = None
        self._ssq_bufs = None
        self._packed = None  # lazily built packed copies of the decoder weights (csrc/gemm_stream.cu)
        # tp > 1: all-reduce inside the persistent kernel over NVLink peer memory (0 = per-op kernels + NCCL all-reduce)
        self.tp_mega_enabled = os.environ.get("OMCHAT_B200_TP_MEGA", "1") != "0"
        self._xchg = {}  # batch -> lib.PeerExchange (tensor-parallel persistent decode kernel)
        # tp > 1, batched step: all-reduce fused into the o_proj / down_proj epilogues over NVLink peer memory
        # (csrc/gemm_stream.cu; 0 = NCCL all-reduce between the kernels)
        self.tp_stream_fused = os.environ.get("OMCHAT_B200_TP_STREAM_FUSED", "1") != "0"
        self._stream_xchg = None  # lib.PeerExchange of the batched step (or a test's emulation object with .ptrs)
        self._mega_epoch = 1  # shared by every plan that uses the peer exchange buffers: identical on all ranks
        self._dec = {}  # decode state per batch size
        self._caches = {}  # reusable caches for generate(), keyed by (n_seq, capacity)

    # ------------------------------------------------------------------------------------------------ helpers
    def new_cache(self, n_seq: int, max_ctx: int, shuffle_pages: bool = True) -> PagedKVCache:
        return PagedKVCache(len(self.w.layers), n_seq, max_ctx, self.Hkv, self.cfg.kv_page_size, self.device,
                            shuffle_pages=shuffle_pages)

    def acquire_cache(self, n_seq: int, max_ctx: int) -> PagedKVCache:
        """A cache owned by the decoder and reused across generate() calls of the same shape (so that the captured
        decode graph, which bakes the pool / block-table pointers in, is reused too). Pages are simply overwritten."""
        cap = ((max_ctx + 255) // 256) * 256
        key = (n_seq, cap)
        c = self._caches.get(key)
        if c is None:
            if len(self._caches) >= 4:  # bound the memory held by idle caches
                self._caches.pop(next(iter(self._caches)))
            c = self.new_cache(n_seq, cap)
            self._caches[key] = c
        c.host_lens = [0] * n_seq
        c.ctx_lens.zero_()
        return c

    def _all_reduce(self, t: torch.Tensor):
        if self.tp.size > 1:
            torch.distributed.all_reduce(t, group=self.tp.group)

    def _folded_prefill(self):
        """(q|k|v * input_layernorm, gate|up * post_attention_layernorm) per layer, row-major for the prefill GEMM; built once."""
        if self._folded_w is None:
            from .weights import fold_norm
            self._folded_w = [(fold_norm(l.qkv_w, l.ln1), fold_norm(l.gate_up_w, l.ln2)) for l in self.w.layers]
        return self._folded_w

    def _prefill_ssq(self, rows: int, device):
        if self._ssq_bufs is None or self._ssq_bufs[0].rows < rows:
            self._ssq_bufs = (lib.RowSsq(rows, device), lib.RowSsq(rows, device))
        return self._ssq_bufs

    def _row_parallel(self, x, w, h, use_gemv: bool, ssq_out=None):
        """h <- h + x @ w^T summed over TP ranks (o_proj / down_proj; modeling_qwen2.py:245,296,303). ssq_out: leave the
        rows' sums of squares for the folded RMSNorm of the next GEMM."""
        if self.tp.size == 1:
            if use_gemv:
                lib.gemv(x, w, out=h, res=h, epi=lib.EPI_RES)
            else:
                lib.gemm(x, w, out=h, res=h, epi=lib.EPI_RES, ssq_out=ssq_out)
            return
        # TP: rank 0 folds the residual into its partial product, then the all-reduce yields the new residual stream
        fold = self.tp.rank == 0
        if use_gemv:
            lib.gemv(x, w, out=h, res=h if fold else None, epi=lib.EPI_RES if fold else lib.EPI_NONE)
        else:
            lib.gemm(x, w, out=h, res=h if fold else None, epi=lib.EPI_RES if fold else lib.EPI_NONE)
        self._all_reduce(h)
        if ssq_out is not None:
            ssq_out.from_rows(h)

    @staticmethod
    def _gemv_ok(B: int, K: int) -> bool:
        return B <= GEMV_MAX_B and B * K * 2 <= GEMV_MAX_SMEM

    def release(self):
        """Drop what references the process group / peer mappings: captured decode graphs (they hold NCCL work under tensor
        parallelism), persistent-kernel plans, CUDA-IPC exchange buffers. Call on every rank before
        torch.distributed.destroy_process_group() - tearing the communicator down under live graphs hangs."""
        for st in self._dec.values():
            st.graphs.clear()
            st.plans.clear()
        self._dec.clear()
        for px in self._xchg.values():
            px.close()
        self._xchg.clear()
        if self._stream_xchg is not None:
            self._stream_xchg.close()
            self._stream_xchg = None
        self._caches.clear()

    # ------------------------------------------------------------------------------------------------ packed weights
    class _Packed:
        pass

    def packed_weights(self):
        """The decoder's matrices re-laid for the weight-streaming GEMM (built once, on the first batched decode step):
        q|k|v and gate|up carry the RMSNorm weight in front of them, lm_head the final norm (modeling_qwen2.py:258-263
        folded into :219-221, :46-48, :470-472). Costs one more copy of the decoder weights in HBM (14 GB of 180)."""
        if self._packed is None:
            P = Qwen2Decoder._Packed()
            P.layers = []
            for l in self.w.layers:
                e = Qwen2Decoder._Packed()
                e.qkv = lib.PackedWeight(l.qkv_w, col_scale=l.ln1)
                e.o = lib.PackedWeight(l.o_w)
                e.gate_up = lib.PackedWeight(l.gate_up_w, col_scale=l.ln2)
                e.down = lib.PackedWeight(l.down_w)
                P.layers.append(e)
            P.lm_head = lib.PackedWeight(self.w.lm_head, col_scale=self.w.norm)
            self._packed = P
        return self._packed

    def use_stream(self, B: int) -> bool:
        """Batches stream_min_b..64 decode on the weight-streaming GEMMs; smaller ones on the persistent kernel (or, with
        that switched off, on the per-op GEMV kernels that serve as its cross-check)."""
        return self.stream_enabled and self.stream_min_b <= B <= 64

    # ------------------------------------------------------------------------------------------------ prefill
    @torch.no_grad()
    def prefill(self, embeds: torch.Tensor, pos_ids: torch.Tensor, seq_ids: torch.Tensor, offsets: Sequence[int],
                cache: PagedKVCache, logits: str = "last", collect_hidden: bool = False,
                slots: Optional[Sequence[int]] = None):
        """embeds [T, C] bf16 packed; pos_ids/seq_ids int32 [T]; offsets: host list of n_seq+1 row offsets.
        Fills the cache for every sequence (fresh prefill from position 0) and returns fp32 logits:
        'last' -> [n_seq, V_local] of each sequence's final token, 'all' -> [T, V_local], 'none' -> None.
        slots: cache rows the packed sequences go to (seq_ids must already hold these row numbers) when only SOME rows of
        the cache are (re)filled - continuous batching; the other rows keep their state."""
        n_seq = len(offsets) - 1
        T = offsets[-1]
        assert embeds.shape[0] >= T and (n_seq == cache.n_seq if slots is None else len(slots) == n_seq)
        lens = [offsets[i + 1] - offsets[i] for i in range(n_seq)]
        assert max(lens) <= cache.capacity, "KV cache too small for this prefill"
        dev = embeds.device
        cu = torch.tensor(list(offsets), dtype=torch.int32).to(dev, non_blocking=True)
        h = embeds[:T].clone() if self.tp.size > 1 else embeds[:T]  # residual stream, updated in place
        C, Hq, Hkv = self.C, self.Hq, self.Hkv
        qw = (Hq + 2 * Hkv) * 128
        xn = torch.empty(T, C, device=dev, dtype=torch.bfloat16)
        qkv = torch.empty(T, qw, device=dev, dtype=torch.bfloat16)
        attn = torch.empty(T, Hq * 128, device=dev, dtype=torch.bfloat16)
        act = torch.empty(T, self.I_local, device=dev, dtype=torch.bfloat16)
        hiddens = [h.clone()] if collect_hidden else None
        max_len = max(lens)
        fold = self.fold_norms
        if fold:
            # input / post-attention RMSNorm folded into the qkv / gate|up GEMMs (omc_gemm_bf16_norm): o_proj / down_proj leave
            # the rows' sums of squares behind (under tensor parallelism: a row pass after the all-reduce)
            folded = self._folded_prefill()
            ssq_a, ssq_b = self._prefill_ssq(T, dev)
            ssq_a.from_rows(h)
        for li, l in enumerate(self.w.layers):
            if fold:
                lib.gemm(h, folded[li][0], out=qkv, bias=l.qkv_b, ssq_in=ssq_a, norm_dim=C, eps=self.eps)
            else:
                lib.rmsnorm(h, l.ln1, self.eps, out=xn)
                lib.gemm(xn, l.qkv_w, out=qkv, bias=l.qkv_b)
            lib.rope_kv_store(qkv, pos_ids, seq_ids, Hq, Hkv, self.inv_freq, cache.pool[li], cache.block_table,
                              cache.page_size)
            lib.attention(qkv[:, :Hq * 128], qkv[:, Hq * 128:(Hq + Hkv) * 128], qkv[:, (Hq + Hkv) * 128:], attn, cu,
                          max_len, Hq, Hkv, True, self.scale)
            self._row_parallel(attn, l.o_w, h, use_gemv=False, ssq_out=ssq_b if fold else None)
            if fold:
                lib.gemm(h, folded[li][1], out=act, epi=lib.EPI_SWIGLU, ssq_in=ssq_b, norm_dim=C, eps=self.eps)
                self._row_parallel(act, l.down_w, h, use_gemv=False, ssq_out=ssq_a)
            else:
                self._mlp_rows(li, l, h, xn, act)
            if collect_hidden:
                hiddens.append(h.clone())
        if slots is None:
            cache.host_lens = list(lens)
            cache.ctx_lens.copy_(torch.tensor(lens, dtype=torch.int32), non_blocking=True)
        else:
            for s_, n_ in zip(slots, lens):
                cache.host_lens[s_] = n_
            cache.ctx_lens.copy_(torch.tensor(cache.host_lens, dtype=torch.int32), non_blocking=True)
        out = None
        if logits == "last":
            last_rows = torch.tensor([offsets[i + 1] - 1 for i in range(n_seq)], dtype=torch.int64).to(dev)
            hl = h.index_select(0, last_rows)
            out = self.lm_head(hl)
        elif logits == "all":
            out = self.lm_head(h)
        return (out, hiddens) if collect_hidden else out

    # the MLP half of a layer on the per-op paths (overridden by the mixture-of-experts decoder, model/moe.py)
    def _mlp_rows(self, li: int, l, h, xn, act):
        """prefill: h += down(silu(gate(norm(h))) * up(norm(h)))   (modeling_qwen2.py:46-48,300-303)"""
        lib.rmsnorm(h, l.ln2, self.eps, out=xn)
        lib.gemm(xn, l.gate_up_w, out=act, epi=lib.EPI_SWIGLU)
        self._row_parallel(act, l.down_w, h, use_gemv=False)

    def _mlp_step(self, li: int, l, h, st, gv_c: bool, gv_i: bool):
        """decode step (B rows): same, on the GEMV kernels (RMSNorm fused) when the batch is small enough"""
        if gv_c:
            lib.gemv(h, l.gate_up_w, out=st.act, norm_w=l.ln2, eps=self.eps, epi=lib.EPI_SWIGLU)
        else:
            lib.rmsnorm(h, l.ln2, self.eps, out=st.xn)
            lib.gemm(st.xn, l.gate_up_w, out=st.act, epi=lib.EPI_SWIGLU)
        self._row_parallel(st.act, l.down_w, h, use_gemv=gv_i)

    @torch.no_grad()
    def lm_head(self, h: torch.Tensor, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """final RMSNorm + lm_head (modeling_qwen2.py:411,470-472) -> fp32 logits [rows, V_local]."""
        rows = h.shape[0]
        if self._gemv_ok(rows, self.C):
            return lib.gemv(h, self.w.lm_head, out=out, norm_w=self.w.norm, eps=self.eps, out_f32=True)
        xn = lib.rmsnorm(h, self.w.norm, self.eps)
        return lib.gemm(xn, self.w.lm_head, out=out, out_f32=True)

    # ------------------------------------------------------------------------------------------------ decode
    class _DecodeState:
        pass

    def _decode_state(self, B: int, max_ctx: int) -> "_DecodeState":
        key = (B, self.Hq, self.Hkv)
        st = self._dec.get(key)
        if st is not None and st.max_ctx >= max_ctx:
            return st
        dev = self.device
        st = Qwen2Decoder._DecodeState()
        st.B, st.max_ctx = B, max_ctx
        st.tokens = torch.zeros(B, device=dev, dtype=torch.int64)
        st.h = torch.empty(B, self.C, device=dev, dtype=torch.bfloat16)
        st.xn = torch.empty(B, self.C, device=dev, dtype=torch.bfloat16)
        st.qkv = torch.empty(B, (self.Hq + 2 * self.Hkv) * 128, device=dev, dtype=torch.bfloat16)
        st.attn = torch.empty(B, self.Hq * 128, device=dev, dtype=torch.bfloat16)
        st.act = torch.empty(B, self.I_local, device=dev, dtype=torch.bfloat16)
        st.logits = torch.empty(B, self.V_local, device=dev, dtype=torch.float32)
        st.arg_ws = torch.empty(128 * B, device=dev, dtype=torch.float32)
        st.ssq_a = torch.zeros(lib.ssq_parts(self.C) * 64, device=dev, dtype=torch.float32)  # row sums of squares entering a layer
        st.ssq_b = torch.zeros(lib.ssq_parts(self.C) * 64, device=dev, dtype=torch.float32)  # ... entering its MLP
        st.splits = lib.decode_attn_splits(B, self.Hkv, max_ctx)
        st.attn_ws = lib.decode_attn_workspace(B, self.Hq, self.Hkv, st.splits, dev)
        if self.tp.size > 1:
            st.loc_val = torch.empty(B, device=dev, dtype=torch.float32)
            st.loc_idx = torch.empty(B, device=dev, dtype=torch.int64)
            st.all_val = torch.empty(self.tp.size, B, device=dev, dtype=torch.float32)
            st.all_idx = torch.empty(self.tp.size, B, device=dev, dtype=torch.int64)
        st.graphs = {}  # id(cache) -> (CUDAGraph, kernel launches per replay, cache)
        st.hist = torch.zeros(MEGA_HIST, B, device=dev, dtype=torch.int64)
        st.hist_pos = torch.zeros(1, device=dev, dtype=torch.int32)
        st.plans = {}  # id(cache) -> (lib.DecodePlan, cache): the persistent decode kernel bakes the cache pointers in
        self._dec[key] = st
        return st

    def use_mega(self, B: int) -> bool:
        """Small-batch decode runs as ONE persistent cooperative kernel per token (csrc/decode_mega.cu)."""
        return (self.mega_enabled and B <= MEGA_MAX_B and B < self.stream_min_b and self.C <= 4096 and len(self.w.layers) <= 32
                and (self.tp.size == 1 or (self.tp.size <= 8 and self.tp_mega_enabled)))

    def _peer_exchange(self, B: int):
        """The per-rank exchange buffers of the tensor-parallel decode kernel (collective: every rank must call this at
        the same point — it does, the first decode step of a batch size)."""
        px = self._xchg.get(B)
        if px is None:
            px = lib.PeerExchange(lib.decode_xchg_bytes(B, self.C, self.tp.size), self.tp.rank, self.tp.size, self.tp.group)
            self._xchg[B] = px
        return px

    def _rope_table(self, positions: int) -> torch.Tensor:
        """(cos, sin)(pos * inv_freq) in fp32 for pos < positions, [positions, 64, 2] — Qwen2RotaryEmbedding.forward
        (modeling_qwen2.py:102-113) evaluated once on the device instead of per step inside the kernel."""
        t = getattr(self, "_rope_cs", None)
        if t is None or t.shape[0] < positions:
            n = max(positions, 2048)
            ang = torch.arange(n, device=self.device, dtype=torch.float32)[:, None] * self.inv_freq[None, :]
            t = torch.stack([ang.cos(), ang.sin()], dim=-1).contiguous()
            self._rope_cs = t
        return t

    def _mega_plan(self, st, cache: PagedKVCache, xchg_ptrs=None, grid=None):
        """xchg_ptrs / grid: overrides for tests that emulate several ranks inside one process (default: the peer
        exchange buffers of the process group, one CTA per SM)."""
        ent = st.plans.get(id(cache))
        if ent is not None and ent[1] is cache:
            return ent[0]
        if len(st.plans) >= 4:
            st.plans.pop(next(iter(st.plans)))
        plan = lib.DecodePlan(
            layers=self.w.layers, embed=self.w.embed, final_norm=self.w.norm, lm_head=self.w.lm_head,
            rope_cs=self._rope_table(cache.capacity), cfg_dims=(self.C, self.Hq, self.Hkv, self.I_local, self.V_local),
            kv_pool=cache.pool, block_table=cache.block_table, ctx_lens=cache.ctx_lens, tokens=st.tokens,
            token_hist=st.hist, hist_pos=st.hist_pos, h=st.h, qkv=st.qkv, attn=st.attn, act=st.act, logits=st.logits,
            page_size=cache.page_size, eps=self.eps, scale=self.scale,
            vocab_offset=self.tp.rank * self.V_local if self.tp.size > 1 else 0, tp_rank=self.tp.rank, tp_size=self.tp.size,
            xchg_ptrs=(xchg_ptrs or self._peer_exchange(st.B).ptrs) if self.tp.size > 1 else None, grid=grid)
        st.plans[id(cache)] = (plan, cache)
        return plan

    def _mega_step(self, plan):
        if self.tp.size == 1:
            plan.step()
        else:
            plan.step(self._mega_epoch)
            self._mega_epoch += 1

    def _decode_body(self, st, cache: PagedKVCache, sample: bool = True):
        """One decode step for st.tokens (the tokens generated last step): embeds them, runs the 28 layers against the
        paged cache (appending their K/V), computes logits and, if `sample`, overwrites st.tokens with the greedy next
        tokens. Only device work, no host sync: capturable in a CUDA graph."""
        B = st.B
        if self.use_mega(B):
            # embed + all layers + lm_head + argmax + ctx_lens += 1 in one launch (always samples into st.tokens)
            self._mega_step(self._mega_plan(st, cache))
            return
        if self.use_stream(B):
            self._decode_body_stream(st, cache)
            if sample:
                self._greedy(st)
            return
        cache.ctx_lens.add_(1)  # context length INCLUDING the token being processed
        lib.embed_lookup(st.tokens, self.w.embed, out=st.h)
        h = st.h
        gv_c = self._gemv_ok(B, self.C)
        gv_a = self._gemv_ok(B, self.Hq * 128)
        gv_i = self._gemv_ok(B, self.I_local)
        for li, l in enumerate(self.w.layers):
            if gv_c:
                lib.gemv(h, l.qkv_w, out=st.qkv, norm_w=l.ln1, eps=self.eps, bias=l.qkv_b)
            else:
                lib.rmsnorm(h, l.ln1, self.eps, out=st.xn)
                lib.gemm(st.xn, l.qkv_w, out=st.qkv, bias=l.qkv_b)
            lib.paged_decode_attn(st.qkv, self.inv_freq, cache.pool[li], cache.block_table, cache.page_size,
                                  cache.ctx_lens, self.Hq, self.Hkv, st.splits, self.scale, st.attn, st.attn_ws)
            self._row_parallel(st.attn, l.o_w, h, use_gemv=gv_a)
            self._mlp_step(li, l, h, st, gv_c, gv_i)
        self.lm_head(h, out=st.logits)
        if sample:
            self._greedy(st)

    def stream_exchange(self):
        """Exchange buffers of the batched step's fused all-reduce (collective: every rank calls this at the same point -
        the first batched decode step)."""
        if self._stream_xchg is None:
            self._stream_xchg = lib.PeerExchange(lib.gemm_stream_xchg_bytes(), self.tp.rank, self.tp.size, self.tp.group)
        return self._stream_xchg

    def _decode_body_stream(self, st, cache: PagedKVCache):
        """Batched decode step on the weight-streaming GEMMs: 5 kernels per layer, each launched as a programmatic
        dependent of the previous one; no stand-alone RMSNorm (folded into the GEMMs that follow it). Under tensor
        parallelism the two all-reduces per layer happen inside the o_proj / down_proj epilogues over NVLink."""
        if self.tp.size > 1 and self.tp_stream_fused:
            return self._decode_body_stream_tp(st, cache)
        P = self.packed_weights()
        C, eps = self.C, self.eps
        tp = self.tp.size > 1
        parts = 1 if tp else lib.ssq_parts(C)
        cache.ctx_lens.add_(1)  # context length INCLUDING the token being processed
        lib.embed_lookup(st.tokens, self.w.embed, out=st.h)
        lib.row_ssq(st.h, st.ssq_a, parts=parts, pdl=False)
        h = st.h
        fold = (not tp) or self.tp.rank == 0  # under TP rank 0 alone adds the residual; the all-reduce completes the sum
        for li, (l, p) in enumerate(zip(self.w.layers, P.layers)):
            lib.gemm_stream(h, p.qkv, out=st.qkv, bias=l.qkv_b, ssq_in=st.ssq_a, ssq_in_parts=parts, norm_dim=C, eps=eps)
            lib.paged_decode_attn(st.qkv, self.inv_freq, cache.pool[li], cache.block_table, cache.page_size,
                                  cache.ctx_lens, self.Hq, self.Hkv, st.splits, self.scale, st.attn, st.attn_ws)
            lib.gemm_stream(st.attn, p.o, out=h, res=h if fold else None, epi=lib.EPI_RES if fold else lib.EPI_NONE,
                            ssq_out=None if tp else st.ssq_b)
            if tp:
                self._all_reduce(h)
                lib.row_ssq(h, st.ssq_b, parts=1, pdl=False)
            lib.gemm_stream(h, p.gate_up, out=st.act, epi=lib.EPI_SWIGLU, ssq_in=st.ssq_b, ssq_in_parts=parts, norm_dim=C,
                            eps=eps)
            lib.gemm_stream(st.act, p.down, out=h, res=h if fold else None, epi=lib.EPI_RES if fold else lib.EPI_NONE,
                            ssq_out=None if tp else st.ssq_a)
            if tp:
                self._all_reduce(h)
                lib.row_ssq(h, st.ssq_a, parts=1, pdl=False)
        lib.gemm_stream(h, P.lm_head, out=st.logits, out_f32=True, ssq_in=st.ssq_a, ssq_in_parts=parts, norm_dim=C, eps=eps)

    def _decode_body_stream_tp(self, st, cache: PagedKVCache):
        P = self.packed_weights()
        C, eps = self.C, self.eps
        parts = lib.ssq_parts(C)
        ptrs = self.stream_exchange().ptrs
        x_o, x_down = lib.tp_xchg(ptrs, self.tp.rank, 0), lib.tp_xchg(ptrs, self.tp.rank, 1)
        cache.ctx_lens.add_(1)
        lib.embed_lookup(st.tokens, self.w.embed, out=st.h)
        lib.row_ssq(st.h, st.ssq_a, parts=parts, pdl=False)
        h = st.h
        for li, (l, p) in enumerate(zip(self.w.layers, P.layers)):
            lib.gemm_stream(h, p.qkv, out=st.qkv, bias=l.qkv_b, ssq_in=st.ssq_a, ssq_in_parts=parts, norm_dim=C, eps=eps)
            lib.paged_decode_attn(st.qkv, self.inv_freq, cache.pool[li], cache.block_table, cache.page_size,
                                  cache.ctx_lens, self.Hq, self.Hkv, st.splits, self.scale, st.attn, st.attn_ws)
            lib.gemm_stream(st.attn, p.o, out=h, res=h, epi=lib.EPI_RES, ssq_out=st.ssq_b, tp=x_o)
            lib.gemm_stream(h, p.gate_up, out=st.act, epi=lib.EPI_SWIGLU, ssq_in=st.ssq_b, ssq_in_parts=parts, norm_dim=C,
                            eps=eps)
            lib.gemm_stream(st.act, p.down, out=h, res=h, epi=lib.EPI_RES, ssq_out=st.ssq_a, tp=x_down)
        lib.gemm_stream(h, P.lm_head, out=st.logits, out_f32=True, ssq_in=st.ssq_a, ssq_in_parts=parts, norm_dim=C, eps=eps)

    def _greedy(self, st):
        """HF GenerationMixin greedy argmax (cli.py:60-70); vocab-parallel under TP."""
        if self.tp.size == 1:
            lib.argmax(st.logits, out=st.tokens, workspace=st.arg_ws)
            return
        lib.argmax(st.logits, out=st.loc_idx, workspace=st.arg_ws)
        torch.gather(st.logits, 1, st.loc_idx.view(-1, 1), out=st.loc_val.view(-1, 1))
        st.loc_idx.add_(self.tp.rank * self.V_local)
        torch.distributed.all_gather_into_tensor(st.all_val, st.loc_val, group=self.tp.group)
        torch.distributed.all_gather_into_tensor(st.all_idx, st.loc_idx, group=self.tp.group)
        best = torch.argmax(st.all_val, dim=0, keepdim=True)  # first max = lowest rank = lowest vocab index
        torch.gather(st.all_idx, 0, best, out=st.tokens.view(1, -1))

    @torch.no_grad()
    def decode_step(self, tokens: torch.Tensor, cache: PagedKVCache, sample: bool = False) -> torch.Tensor:
        """Eager single step (the forward(input_ids=[b,1], past_key_values=cache) path). Returns fp32 logits
        [B, V_local] (a view of a static buffer)."""
        B = tokens.numel()
        need = max(cache.host_lens) + 1
        assert need <= cache.capacity, "KV cache is full"
        st = self._decode_state(B, cache.capacity)
        st.tokens.copy_(tokens.view(-1))
        self._decode_body(st, cache, sample=sample)
        cache.host_lens = [n + 1 for n in cache.host_lens]
        return st.logits

    @torch.no_grad()
    def generate_greedy(self, first_tokens: torch.Tensor, cache: PagedKVCache, steps: int, use_graph: bool = True,
                        on_token=None) -> torch.Tensor:
        """Runs `steps` decode steps starting from first_tokens [B] (already sampled from the prefill logits).
        Returns the [B, steps] tokens produced by those steps. With use_graph the step is captured once per
        (batch, cache) and replayed; `on_token(step, tokens)` (streamer hook) forces eager per-step host access."""
        B = first_tokens.numel()
        assert max(cache.host_lens) + steps <= cache.capacity, "KV cache too small for the requested number of tokens"
        st = self._decode_state(B, cache.capacity)
        st.tokens.copy_(first_tokens.view(-1))
        out = torch.empty(steps, B, device=self.device, dtype=torch.int64)
        if steps <= 0:
            return out.t()
        if self.use_mega(B):
            # one persistent kernel per token; the kernel appends every sampled token to st.hist on the device
            plan = self._mega_plan(st, cache)
            done = 0
            while done < steps:
                n = min(MEGA_HIST, steps - done)
                st.hist_pos.zero_()
                for i in range(n):
                    self._mega_step(plan)
                    if on_token is not None:
                        on_token(done + i, st.tokens)
                out[done:done + n].copy_(st.hist[:n])
                done += n
            cache.host_lens = [x + steps for x in cache.host_lens]
            return out.t()
        if use_graph and id(cache) not in st.graphs:
            # warm up once eagerly (sets kernel attributes, loads modules), then capture
            snapshot = (cache.ctx_lens.clone(), st.tokens.clone())
            s = torch.cuda.Stream()
            s.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s):
                self._decode_body(st, cache)
            torch.cuda.current_stream().wait_stream(s)
            cache.ctx_lens.copy_(snapshot[0])
            st.tokens.copy_(snapshot[1])
            g = torch.cuda.CUDAGraph()
            n0 = lib.launch_count()
            with torch.cuda.graph(g):
                self._decode_body(st, cache)
            n_launch = lib.launch_count() - n0
            lib.add_launches(-n_launch)  # capture does not execute
            if len(st.graphs) >= 4:
                st.graphs.pop(next(iter(st.graphs)))
            st.graphs[id(cache)] = (g, n_launch, cache)
            # the capture pass does not execute; restore is unnecessary, but the warm-up wrote K/V of a junk step at
            # position ctx (overwritten by the real step) - harmless.
        graph, n_launch = st.graphs[id(cache)][:2] if use_graph else (None, 0)
        for i in range(steps):
            if use_graph:
                graph.replay()
                lib.add_launches(n_launch)
            else:
                self._decode_body(st, cache)
            out[i].copy_(st.tokens)
            if on_token is not None:
                on_token(i, st.tokens)
        cache.host_lens = [n + steps for n in cache.host_lens]
        return out.t()
