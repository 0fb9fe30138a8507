"""World-size-2 (gloo, CPU) test of the tensor-parallel host logic: the shards that omchat_b200.model.weights cuts
(tp_plan / shard_llm_layer: fused q|k|v, interleaved gate/up, row-parallel o/down, vocab-parallel lm_head, replicated kv
heads when tp > kv heads) reproduce the unsharded decoder when the row-parallel partial sums are all-reduced and the
(max, index) candidates of the vocab-parallel argmax are gathered — the algebra both the NCCL path and the in-kernel
NVLink exchange of csrc/decode_mega.cu implement. Checker: the fp32 oracle (transformers modeling_qwen2.py restatement).
"""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import omchat_oracle as O  # checker only
from omchat_b200.config import OmChatQwen2Config
from omchat_b200.model.weights import shard_llm_layer, tp_plan


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _full_sd(cfg, seed=0):
    g = torch.Generator().manual_seed(seed)
    H, I, D, V = cfg.hidden_size, cfg.intermediate_size, cfg.head_dim, cfg.vocab_size
    nq, nkv = cfg.num_attention_heads, cfg.num_key_value_heads
    rn = lambda *s, std=0.05, mean=0.0: torch.randn(*s, generator=g) * std + mean  # noqa: E731
    sd = {"model.embed_tokens.weight": rn(V, H, std=0.5), "model.norm.weight": rn(H, std=0.1, mean=1.0),
          "lm_head.weight": rn(V, H)}
    for li in range(cfg.num_hidden_layers):
        p = f"model.layers.{li}."
        sd.update({p + "self_attn.q_proj.weight": rn(nq * D, H), p + "self_attn.q_proj.bias": rn(nq * D, std=0.1),
                   p + "self_attn.k_proj.weight": rn(nkv * D, H), p + "self_attn.k_proj.bias": rn(nkv * D, std=0.1),
                   p + "self_attn.v_proj.weight": rn(nkv * D, H), p + "self_attn.v_proj.bias": rn(nkv * D, std=0.1),
                   p + "self_attn.o_proj.weight": rn(H, nq * D), p + "mlp.gate_proj.weight": rn(I, H),
                   p + "mlp.up_proj.weight": rn(I, H), p + "mlp.down_proj.weight": rn(H, I),
                   p + "input_layernorm.weight": rn(H, std=0.1, mean=1.0),
                   p + "post_attention_layernorm.weight": rn(H, std=0.1, mean=1.0)})
    return sd


def _sharded_forward(sd, cfg, ocfg, rank, size, embeds, pos):
    """One rank's share of Qwen2Model.forward on the fused shard layouts, all-reducing where the kernels do."""
    plan = tp_plan(cfg, rank, size)
    D = cfg.head_dim
    Hq, Hkv = len(plan.q_heads), len(plan.kv_heads)
    h = embeds.clone()
    cos, sin = O.rope_cos_sin(pos, ocfg, h.dtype)
    T = h.shape[1]
    for li in range(cfg.num_hidden_layers):
        p, a = f"model.layers.{li}.", f"model.layers.{li}.self_attn."
        qkv_w, qkv_b, o_w, gu_w, down_w = shard_llm_layer(
            sd[a + "q_proj.weight"], sd[a + "q_proj.bias"], sd[a + "k_proj.weight"], sd[a + "k_proj.bias"],
            sd[a + "v_proj.weight"], sd[a + "v_proj.bias"], sd[a + "o_proj.weight"], sd[p + "mlp.gate_proj.weight"],
            sd[p + "mlp.up_proj.weight"], sd[p + "mlp.down_proj.weight"], plan, D)
        x = O.rms_norm(h, sd[p + "input_layernorm.weight"], cfg.rms_norm_eps)
        qkv = F.linear(x, qkv_w, qkv_b)
        q = qkv[..., :Hq * D].view(1, T, Hq, D).transpose(1, 2)
        k = qkv[..., Hq * D:(Hq + Hkv) * D].view(1, T, Hkv, D).transpose(1, 2)
        v = qkv[..., (Hq + Hkv) * D:].view(1, T, Hkv, D).transpose(1, 2)
        q, k = O.apply_rope(q, k, cos, sin)
        G = Hq // Hkv
        kr = k[:, :, None].expand(1, Hkv, G, T, D).reshape(1, Hq, T, D)
        vr = v[:, :, None].expand(1, Hkv, G, T, D).reshape(1, Hq, T, D)
        w = (q @ kr.transpose(2, 3)) * D ** -0.5
        w = w.masked_fill(~torch.tril(torch.ones(T, T, dtype=torch.bool)), torch.finfo(w.dtype).min)
        ctx = (torch.softmax(w, dim=-1) @ vr).transpose(1, 2).reshape(1, T, Hq * D)
        part = F.linear(ctx, o_w)  # row-parallel partial sum
        dist.all_reduce(part)
        h = h + part
        x = O.rms_norm(h, sd[p + "post_attention_layernorm.weight"], cfg.rms_norm_eps)
        gu = F.linear(x, gu_w)  # rows alternate gate_i, up_i
        part = F.linear(F.silu(gu[..., 0::2]) * gu[..., 1::2], down_w)
        dist.all_reduce(part)
        h = h + part
    hn = O.rms_norm(h, sd["model.norm.weight"], cfg.rms_norm_eps)
    logits = F.linear(hn, sd["lm_head.weight"][plan.v_lo:plan.v_hi])  # vocab-parallel
    val, idx = logits[0, -1].max(dim=0)
    cand = torch.tensor([float(val), float(idx + plan.v_lo)], dtype=torch.float64)
    allc = [torch.zeros(2, dtype=torch.float64) for _ in range(size)]
    dist.all_gather(allc, cand)
    best = max(allc, key=lambda c: (float(c[0]), -float(c[1])))  # highest value, lowest index on ties
    full_logits = [torch.zeros_like(logits) for _ in range(size)]
    dist.all_gather(full_logits, logits.contiguous())
    return torch.cat(full_logits, dim=-1), int(best[1])


def _worker(rank, size, port, heads, kv_heads, ret):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=size)
    try:
        torch.manual_seed(0)
        torch.set_num_threads(1)
        cfg = OmChatQwen2Config(vocab_size=640, hidden_size=heads * 128, intermediate_size=1000, num_hidden_layers=2,
                                num_attention_heads=heads, num_key_value_heads=kv_heads, mm_vision_tower=None)
        ocfg = O.OracleConfig(hidden=cfg.hidden_size, heads=heads, kv_heads=kv_heads, inter=1000, layers=2, vocab=640,
                              rope_theta=cfg.rope_theta)
        sd = _full_sd(cfg)
        ids = torch.randint(0, 640, (1, 19), generator=torch.Generator().manual_seed(3))
        emb = sd["model.embed_tokens.weight"][ids[0]][None]
        pos = torch.arange(19)[None]
        want, _ = O.qwen2_forward(emb, pos, sd, ocfg)
        got, tok = _sharded_forward(sd, cfg, ocfg, rank, size, emb, pos)
        err = (got - want).abs().max().item() / want.abs().max().item()
        ok = err < 1e-4 and tok == int(want[0, -1].argmax())
        ret[rank] = (ok, err, tok, int(want[0, -1].argmax()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("heads,kv_heads", [(4, 2), (4, 1)])  # (4,1): tp 2 > kv heads -> the kv head is replicated
def test_tp2_shards_reproduce_unsharded_decoder(heads, kv_heads):
    size, port = 2, _free_port()
    ret = mp.get_context("spawn").Manager().dict()
    mp.spawn(_worker, args=(size, port, heads, kv_heads, ret), nprocs=size, join=True)
    assert len(ret) == size
    for r in range(size):
        ok, err, tok, want = ret[r]
        assert ok, (r, err, tok, want)
    assert ret[0][2] == ret[1][2]
