"""Qwen2-MoE language model variant: OmChatQwen2MoeForCausalLM (omchat/model/language_model/omchat_qwen2_moe.py:28-117), i.e.
the same multimodal glue over transformers' Qwen2MoeForCausalLM instead of Qwen2ForCausalLM.

Attention, RMSNorm, RoPE, the paged cache and lm_head are the dense decoder's kernels; only the MLP half of a *sparse* layer
differs (transformers modeling_qwen2_moe.py:295-374). It runs as: omc_moe_route (RMSNorm, router softmax, top-k, histogram) ->
omc_moe_plan -> omc_moe_scatter (rows sorted by expert into 128-row tiles) -> two grouped tcgen05 GEMMs over the stacked expert
matrices (gate|up with the SwiGLU epilogue, down) -> the shared expert's two GEMMs -> omc_moe_combine (weighted sum + sigmoid-gated
shared expert + residual). Everything stays on the device (no host look at the routing), so decode steps capture into a CUDA
graph like the dense model's. Layers listed in mlp_only_layers (or skipped by decoder_sparse_step) keep the dense MLP.

Decode steps of 1..64 rows run the attention projections, the shared expert and lm_head on the weight-streaming GEMMs of
csrc/gemm_stream.cu (RMSNorms folded in) around the routed block. Not built for this variant: tensor parallelism (expert
parallelism is the natural sharding) and the persistent decode kernel.
"""
from __future__ import annotations

import os
from typing import Optional

import torch

from .. import lib
from ..config import OmChatQwen2MoeConfig
from .decoder import Qwen2Decoder, TPInfo
from .omchat import OmChatQwen2ForCausalLM, OmChatQwen2Model
from .weights import LlmW


class Qwen2MoeDecoder(Qwen2Decoder):
    def __init__(self, cfg: OmChatQwen2MoeConfig, w: LlmW, tp: Optional[TPInfo] = None):
        super().__init__(cfg, w, tp)
        if self.tp.size != 1:
            raise NotImplementedError("tensor parallelism is not built for the Qwen2-MoE variant")
        self.mega_enabled = False
        self.fold_norms = False
        # decode steps of 1..64 rows: attention projections, the shared expert and lm_head on the weight-streaming GEMMs
        # (csrc/gemm_stream.cu, RMSNorms folded in); the routed experts on the grouped GEMM. OMCHAT_B200_NO_STREAM=1: per-op path
        self.stream_min_b = 1
        self.overlap_shared = os.environ.get("OMCHAT_B200_MOE_OVERLAP", "1") != "0"
        self._side_stream = torch.cuda.Stream(device=self.device)
        self._moe_ws = {}
        self._rcat = {}

    def _workspace(self, T: int) -> lib.MoeWorkspace:
        """One workspace per power-of-two token bucket, shared by all layers (allocated outside graph capture: the first,
        uncaptured step of generate_greedy creates it)."""
        bucket = 1 << max(T - 1, 0).bit_length()
        ws = self._moe_ws.get(bucket)
        if ws is None:
            c = self.cfg
            ws = lib.MoeWorkspace(bucket, self.C, c.num_experts, c.num_experts_per_tok, c.moe_intermediate_size,
                                  c.shared_expert_intermediate_size, self.device)
            self._moe_ws[bucket] = ws
        return ws

    def _router_cat(self, m):
        """Router rows stacked with the shared expert's gate row for the tensor-core logits GEMM of prefill-sized blocks."""
        key = m.router_w.data_ptr()
        if key not in self._rcat:
            self._rcat[key] = lib.router_cat(m.router_w, m.shared_gate_w)
        return self._rcat[key]

    def _moe(self, l, h, xn):
        m = l.moe
        # post_attention_layernorm is applied by the router kernel (it needs the normed row anyway) and left in xn
        lib.moe_block(h, xn[:h.shape[0]], self._workspace(h.shape[0]), m.router_w, m.shared_gate_w, m.experts_gate_up,
                      m.experts_down, m.shared_gate_up, m.shared_down, self.cfg.norm_topk_prob, norm_w=l.ln2, eps=self.eps,
                      router_cat_w=self._router_cat(m) if h.shape[0] >= lib.ROUTER_GEMM_MIN_T else None)

    def _mlp_rows(self, li, l, h, xn, act):
        if l.moe is None:
            return super()._mlp_rows(li, l, h, xn, act)
        self._moe(l, h, xn)

    def _mlp_step(self, li, l, h, st, gv_c, gv_i):
        if l.moe is None:
            return super()._mlp_step(li, l, h, st, gv_c, gv_i)
        self._moe(l, h, st.xn)

    def packed_weights(self):
        """As the dense decoder's, with a sparse layer's SHARED expert in the gate_up / down places (the routed experts stay in
        the grouped GEMM's stacked layout)."""
        if self._packed is None:
            P = Qwen2Decoder._Packed()
            P.layers = []
            for l in self.w.layers:
                e = Qwen2Decoder._Packed()
                e.qkv = lib.PackedWeight(l.qkv_w, col_scale=l.ln1)
                e.o = lib.PackedWeight(l.o_w)
                gu, dn = (l.gate_up_w, l.down_w) if l.moe is None else (l.moe.shared_gate_up, l.moe.shared_down)
                e.gate_up = lib.PackedWeight(gu, col_scale=l.ln2)
                e.down = lib.PackedWeight(dn)
                P.layers.append(e)
            P.lm_head = lib.PackedWeight(self.w.lm_head, col_scale=self.w.norm)
            self._packed = P
        return self._packed

    def _decode_body_stream(self, st, cache):
        """Decode step (1..64 rows): per layer qkv -> paged attention -> o_proj (+residual, leaves the rows' sums of squares) ->
        [sparse layer: shared expert gate|up, down on the streaming GEMMs -> route (RMSNorm fused) -> plan -> scatter ->
        grouped gate|up, down -> combine -> sums of squares] or [dense layer: gate|up, down]."""
        P = self.packed_weights()
        C, eps, B = self.C, self.eps, st.B
        parts = lib.ssq_parts(C)
        cache.ctx_lens.add_(1)
        lib.embed_lookup(st.tokens, self.w.embed, out=st.h)
        lib.row_ssq(st.h, st.ssq_a, parts=parts, pdl=False)
        h = st.h
        for li, (l, p) in enumerate(zip(self.w.layers, P.layers)):
            lib.gemm_stream(h, p.qkv, out=st.qkv, bias=l.qkv_b, ssq_in=st.ssq_a, ssq_in_parts=parts, norm_dim=C, eps=eps)
            lib.paged_decode_attn(st.qkv, self.inv_freq, cache.pool[li], cache.block_table, cache.page_size,
                                  cache.ctx_lens, self.Hq, self.Hkv, st.splits, self.scale, st.attn, st.attn_ws)
            lib.gemm_stream(st.attn, p.o, out=h, res=h, epi=lib.EPI_RES, ssq_out=st.ssq_b)
            if l.moe is None:
                lib.gemm_stream(h, p.gate_up, out=st.act, epi=lib.EPI_SWIGLU, ssq_in=st.ssq_b, ssq_in_parts=parts, norm_dim=C,
                                eps=eps)
                lib.gemm_stream(st.act, p.down, out=h, res=h, epi=lib.EPI_RES, ssq_out=st.ssq_a)
                continue
            m, ws = l.moe, self._workspace(B)
            # the shared expert (two streaming GEMMs) and the routed block (route -> plan/scatter -> two grouped GEMMs) both
            # read h and are independent until the combine: run them on two streams (a fork / join inside the captured graph)
            main = torch.cuda.current_stream()
            side = self._side_stream if self.overlap_shared else main
            if side is not main:
                side.wait_stream(main)
            with torch.cuda.stream(side):
                lib.gemm_stream(h, p.gate_up, out=ws.shared_act[:B], epi=lib.EPI_SWIGLU, ssq_in=st.ssq_b, ssq_in_parts=parts,
                                norm_dim=C, eps=eps)
                shared_y = lib.gemm_stream(ws.shared_act[:B], p.down, out=ws.shared_y[:B])
            lib.moe_block(h, st.xn, ws, m.router_w, m.shared_gate_w, m.experts_gate_up, m.experts_down, None, None,
                          self.cfg.norm_topk_prob, norm_w=l.ln2, eps=eps, defer_combine=True)
            if side is not main:
                main.wait_stream(side)
            lib.moe_combine(h, ws, shared_y, st.ssq_a, parts)
        lib.gemm_stream(h, P.lm_head, out=st.logits, out_f32=True, ssq_in=st.ssq_a, ssq_in_parts=parts, norm_dim=C, eps=eps)

    def release(self):
        super().release()
        self._moe_ws.clear()
        self._rcat.clear()


class OmChatQwen2MoeModel(OmChatQwen2Model):
    """OmChatQwen2MoeModel(OmChatMetaModel, Qwen2MoeModel) (omchat_qwen2_moe.py:20-24)."""
    decoder_class = Qwen2MoeDecoder


class OmChatQwen2MoeForCausalLM(OmChatQwen2ForCausalLM):
    """Same surface as OmChatQwen2ForCausalLM - forward(images=...), generate(), prepare_inputs_for_generation - over the
    mixture-of-experts decoder (omchat_qwen2_moe.py:27-114)."""
    config_class = OmChatQwen2MoeConfig
    model_class = OmChatQwen2MoeModel
