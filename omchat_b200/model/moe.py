"""Qwen2-MoE language model variant: OmChatQwen2MoeForCausalLM (omchat/model/language_model/omchat_qwen2_moe.py:28-117), i.e.
the same multimodal glue over transformers' Qwen2MoeForCausalLM instead of Qwen2ForCausalLM.

Attention, RMSNorm, RoPE, the paged cache and lm_head are the dense decoder's kernels; only the MLP half of a *sparse* layer
differs (transformers modeling_qwen2_moe.py:295-374). It runs as: omc_moe_route (RMSNorm, router softmax, top-k, histogram) ->
omc_moe_plan -> omc_moe_scatter (rows sorted by expert into 128-row tiles) -> two grouped tcgen05 GEMMs over the stacked expert
matrices (gate|up with the SwiGLU epilogue, down) -> the shared expert's two GEMMs -> omc_moe_combine (weighted sum + sigmoid-gated
shared expert + residual). Everything stays on the device (no host look at the routing), so decode steps capture into a CUDA
graph like the dense model's. Layers listed in mlp_only_layers (or skipped by decoder_sparse_step) keep the dense MLP.

Not built for this variant: tensor parallelism, the persistent decode kernel and the weight-streaming GEMMs (their layer loops
are dense-MLP specific) - every batch size decodes on the per-op path.
"""
from __future__ import annotations

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
        self.stream_enabled = False
        self.fold_norms = False
        self._moe_ws = {}

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

    def _moe(self, l, h, xn):
        m = l.moe
        # post_attention_layernorm is applied by the router kernel (it needs the normed row anyway) and left in xn
        lib.moe_block(h, xn[:h.shape[0]], self._workspace(h.shape[0]), m.router_w, m.shared_gate_w, m.experts_gate_up,
                      m.experts_down, m.shared_gate_up, m.shared_down, self.cfg.norm_topk_prob, norm_w=l.ln2, eps=self.eps)

    def _mlp_rows(self, li, l, h, xn, act):
        if l.moe is None:
            return super()._mlp_rows(li, l, h, xn, act)
        self._moe(l, h, xn)

    def _mlp_step(self, li, l, h, st, gv_c, gv_i):
        if l.moe is None:
            return super()._mlp_step(li, l, h, st, gv_c, gv_i)
        self._moe(l, h, st.xn)

    def release(self):
        super().release()
        self._moe_ws.clear()


class OmChatQwen2MoeModel(OmChatQwen2Model):
    """OmChatQwen2MoeModel(OmChatMetaModel, Qwen2MoeModel) (omchat_qwen2_moe.py:20-24)."""
    decoder_class = Qwen2MoeDecoder


class OmChatQwen2MoeForCausalLM(OmChatQwen2ForCausalLM):
    """Same surface as OmChatQwen2ForCausalLM - forward(images=...), generate(), prepare_inputs_for_generation - over the
    mixture-of-experts decoder (omchat_qwen2_moe.py:27-114)."""
    config_class = OmChatQwen2MoeConfig
    model_class = OmChatQwen2MoeModel
