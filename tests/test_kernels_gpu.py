"""GPU parity tests, kernel by kernel, through the C-ABI (omchat_b200.lib -> libomchat_b200.so) against the CPU oracle
(oracle/omchat_oracle.py) or, for a single linear/softmax op, the same fp32 formula on CPU.

Tolerances (bf16 inputs, fp32 accumulation, one bf16 rounding on output): relative max-abs error <= 2^-7 of the output
scale for GEMM-like ops, cosine >= 0.999 per row; integer/index kernels are bit-exact.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

from oracle import omchat_oracle as O  # noqa: E402  (checker only)


@pytest.fixture(scope="module")
def L():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device")
    from omchat_b200 import lib
    lib.load()
    return lib


def bf(t):
    return t.to(torch.bfloat16)


def assert_close(got, ref, rel=2 ** -7, what=""):
    got, ref = got.float().cpu(), ref.float().cpu()
    assert got.shape == ref.shape, (got.shape, ref.shape)
    assert torch.isfinite(got).all(), f"{what}: non-finite output"
    scale = ref.abs().max().item() + 1e-6
    err = (got - ref).abs().max().item()
    assert err <= rel * scale, f"{what}: max abs err {err:.5g} vs scale {scale:.5g} (rel {err / scale:.4g})"
    if got.dim() == 2 and got.shape[1] >= 8:
        cos = torch.nn.functional.cosine_similarity(got, ref, dim=-1)
        assert cos.min().item() >= 0.999, f"{what}: min row cosine {cos.min().item():.6f}"


def ref_linear(x, w, bias=None):
    return torch.nn.functional.linear(x.float().cpu(), w.float().cpu(), None if bias is None else bias.float().cpu())


# ----------------------------------------------------------------------------------------------- GEMM
GEMM_CFGS = [(256, 1), (128, 1), (256, 2), (192, 2), (160, 2), (128, 2)]


@pytest.mark.parametrize("bn,cg", GEMM_CFGS)
@pytest.mark.parametrize("M,N,K", [(128 * 2, 256 * 3, 64), (256, 768, 512), (1025, 1920, 320), (300, 3840, 1024)])
def test_gemm_plain(L, bn, cg, M, N, K):
    g = torch.Generator().manual_seed(M + N + K)
    x = bf(torch.randn(M, K, generator=g)).cuda()
    w = bf(torch.randn(N, K, generator=g) * 0.1).cuda()
    out = L.gemm(x, w, tile_cfg=bn | (cg << 16))
    torch.cuda.synchronize()
    assert_close(out, ref_linear(x, w), what=f"gemm bn={bn} cg={cg} {M}x{N}x{K}")


@pytest.mark.parametrize("bn,cg", [(256, 1), (256, 2), (160, 2)])
def test_gemm_persistent_many_tiles_few_ctas(L, bn, cg):
    # more tiles than CTAs: exercises the smem ring wrap, both TMEM accumulators and phase flips
    M, N, K = 1280, 1920, 448
    g = torch.Generator().manual_seed(5)
    x = bf(torch.randn(M, K, generator=g)).cuda()
    w = bf(torch.randn(N, K, generator=g) * 0.1).cuda()
    out = L.gemm(x, w, tile_cfg=bn | (cg << 16) | (6 << 20))
    torch.cuda.synchronize()
    assert_close(out, ref_linear(x, w), what="persistent gemm")


@pytest.mark.parametrize("cfg", [0, 256 | (1 << 16), 256 | (2 << 16)])
def test_gemm_epilogues(L, cfg):
    M, N, K = 515, 1280, 384
    g = torch.Generator().manual_seed(7)
    x = bf(torch.randn(M, K, generator=g)).cuda()
    w = bf(torch.randn(N, K, generator=g) * 0.1).cuda()
    bias = bf(torch.randn(N, generator=g)).cuda()
    scale = bf(torch.randn(N, generator=g) * 0.1 + 0.1).cuda()
    res = bf(torch.randn(M, N, generator=g)).cuda()
    lin = ref_linear(x, w, bias)
    assert_close(L.gemm(x, w, bias=bias, tile_cfg=cfg), lin, what="bias")
    assert_close(L.gemm(x, w, bias=bias, epi=L.EPI_GELU, tile_cfg=cfg), torch.nn.functional.gelu(lin), what="gelu")
    want = res.float().cpu() + scale.float().cpu() * lin
    assert_close(L.gemm(x, w, bias=bias, scale=scale, res=res, epi=L.EPI_RES, tile_cfg=cfg), want, what="res")
    # in-place residual (out aliases res), no bias / scale: the decoder's o_proj / down_proj form
    h = res.clone()
    L.gemm(x, w, out=h, res=h, epi=L.EPI_RES, tile_cfg=cfg)
    assert_close(h, res.float().cpu() + ref_linear(x, w), what="res in place")
    assert_close(L.gemm(x, w, out_f32=True, tile_cfg=cfg), ref_linear(x, w), what="f32 out")


def interleave_gate_up(gate, up):
    """[I,K],[I,K] -> [2I,K] with rows alternating gate_i, up_i (the layout EPI_SWIGLU expects)."""
    I, K = gate.shape
    return torch.stack([gate, up], dim=1).reshape(2 * I, K)


@pytest.mark.parametrize("cg", [1, 2])
def test_gemm_swiglu(L, cg):
    M, I, K = 300, 512, 256
    g = torch.Generator().manual_seed(9)
    x = bf(torch.randn(M, K, generator=g)).cuda()
    gate = bf(torch.randn(I, K, generator=g) * 0.1)
    up = bf(torch.randn(I, K, generator=g) * 0.1)
    w = interleave_gate_up(gate, up).cuda()
    out = L.gemm(x, w, epi=L.EPI_SWIGLU, tile_cfg=256 | (cg << 16))
    want = torch.nn.functional.silu(ref_linear(x, gate)) * ref_linear(x, up)
    assert_close(out, want, what="swiglu")


@pytest.mark.parametrize("bn,cg", GEMM_CFGS)
@pytest.mark.parametrize("M,C,N2", [(300, 1920, 768), (1025, 3200, 1280), (77, 256, 512)])
def test_gemm_folded_rmsnorm(L, bn, cg, M, C, N2):
    """RMSNorm folded into the GEMM that follows it (omc_gemm_bf16_norm): an EPI_RES GEMM writes the residual rows and
    their sums of squares (one fp32 partial per N tile, of the bf16 values stored), the next GEMM runs on the raw rows with
    W * g and scales row m by rsqrt(sum / C + eps) in its epilogue. Reference: InternRMSNorm / Qwen2RMSNorm + Linear of the
    oracle in fp32 on the same bf16 inputs (the norm-weight rounding moves from the activations into W': 2^-6)."""
    from omchat_b200.model.weights import fold_norm
    g = torch.Generator().manual_seed(M + C + bn)
    x = bf(torch.randn(M, 512, generator=g)).cuda()
    w1 = bf(torch.randn(C, 512, generator=g) * 0.05).cuda()
    res = bf(torch.randn(M, C, generator=g)).cuda()
    ls = bf(torch.randn(C, generator=g) * 0.05 + 0.1).cuda()
    gw = bf(torch.randn(C, generator=g) * 0.1 + 1.0).cuda()
    w2 = bf(torch.randn(N2, C, generator=g) * 0.05).cuda()
    b2 = bf(torch.randn(N2, generator=g)).cuda()
    cfg = bn | (cg << 16)
    ssq = L.RowSsq(M + 5, "cuda")
    h = res.clone()
    L.gemm(x, w1, out=h, scale=ls, res=h, epi=L.EPI_RES, tile_cfg=cfg, ssq_out=ssq)
    assert ssq.parts == (C + bn - 1) // bn
    assert_close(h, res.float().cpu() + ls.float().cpu() * ref_linear(x, w1), what="res rows")
    got_ssq = ssq.buf[:ssq.parts, :M].sum(0).cpu()
    assert torch.allclose(got_ssq, h.float().pow(2).sum(-1).cpu(), rtol=1e-4), "sums of squares of the rows actually written"
    y = L.gemm(h, fold_norm(w2, gw), bias=b2, epi=L.EPI_GELU, tile_cfg=cfg, ssq_in=ssq, norm_dim=C, eps=1e-6)
    want = torch.nn.functional.gelu(ref_linear(O.rms_norm(h.cpu(), gw.cpu(), 1e-6), w2, b2))
    assert_close(y, want, rel=2 ** -6, what="folded rmsnorm + gelu")
    ssq1 = L.RowSsq(M, "cuda").from_rows(h)
    assert ssq1.parts == 1 and torch.allclose(ssq1.buf[0, :M].cpu(), got_ssq, rtol=1e-4)
    if bn == 256 and N2 % 16 == 0:  # SwiGLU reads the folded norm too (Qwen2 post-attention norm -> gate|up)
        w2f = fold_norm(w2, gw)
        y2 = L.gemm(h, w2f, epi=L.EPI_SWIGLU, tile_cfg=cfg, ssq_in=ssq1, norm_dim=C, eps=1e-6)
        xn = O.rms_norm(h.cpu(), gw.cpu(), 1e-6)
        want2 = torch.nn.functional.silu(ref_linear(xn, w2[0::2])) * ref_linear(xn, w2[1::2])
        assert_close(y2, want2, rel=2 ** -6, what="folded rmsnorm + swiglu")


@pytest.mark.parametrize("M", [1, 9, 16, 17, 32, 33, 64])
@pytest.mark.parametrize("N,K", [(256, 64), (3584, 1792), (1000, 4104), (37888 // 8, 512)])
def test_gemm_skinny(L, M, N, K):
    """Swapped-operand tcgen05 GEMM for M <= 64 (csrc/gemm_skinny.cu): split-K over a thread-block cluster (partials summed
    in rank order through distributed shared memory: bit-identical from call to call) for small N, ragged N / K tails,
    every epilogue."""
    g = torch.Generator().manual_seed(M * 7 + N + K)
    x = bf(torch.randn(M, K, generator=g)).cuda()
    w = bf(torch.randn(N, K, generator=g) * 0.1).cuda()
    bias = bf(torch.randn(N, generator=g)).cuda()
    scale = bf(torch.randn(N, generator=g) * 0.1 + 0.1).cuda()
    res = bf(torch.randn(M, N, generator=g)).cuda()
    lin = ref_linear(x, w, bias)
    assert L.SKINNY_ENABLED and M <= L.SKINNY_MAX_M
    outs = []
    for rep in range(2):
        outs.append(L.gemm(x, w))
        assert_close(outs[-1], ref_linear(x, w), what=f"skinny plain rep {rep}")
    assert torch.equal(outs[0], outs[1]), "the split-K reduction must be deterministic"
    assert_close(L.gemm(x, w, bias=bias, epi=L.EPI_GELU), torch.nn.functional.gelu(lin), what="skinny gelu")
    want = res.float().cpu() + scale.float().cpu() * lin
    assert_close(L.gemm(x, w, bias=bias, scale=scale, res=res, epi=L.EPI_RES), want, what="skinny res")
    h = res.clone()
    L.gemm(x, w, out=h, res=h, epi=L.EPI_RES)
    assert_close(h, res.float().cpu() + ref_linear(x, w), what="skinny res in place")
    assert_close(L.gemm(x, w, out_f32=True), ref_linear(x, w), what="skinny f32 out")
    if N % 16 == 0:
        gate, up = w[0::2].cpu(), w[1::2].cpu()
        wantg = torch.nn.functional.silu(ref_linear(x, gate)) * ref_linear(x, up)
        assert_close(L.gemm(x, w, epi=L.EPI_SWIGLU), wantg, what="skinny swiglu")
    for ws in L._skinny_ws.values():
        assert int(ws.count_nonzero()) == 0, "split-K workspace must be left zeroed"


@pytest.mark.parametrize("M", [1, 9, 16, 17, 32, 33, 64])
@pytest.mark.parametrize("N,K", [(256, 64), (3584, 1792), (1000, 4104), (37888 // 8, 512), (4608, 3584), (152064 // 8, 3584)])
def test_gemm_stream(L, M, N, K):
    """Weight-streaming GEMM of the batched decode step (csrc/gemm_stream.cu): packed 16 KB weight tiles, stream-K over
    2 CTAs per SM with tiles cut across CTAs (partials summed in CTA order: bit-identical from call to call), ragged N / K
    tails, every epilogue, and the folded RMSNorm: y = rstd[m] * (x @ (W * g)^T) with the sums of squares taken from an
    EPI_RES epilogue's partials. Reference: fp32 on the CPU from the same bf16 inputs; the folded-norm case against
    Qwen2RMSNorm + Linear of the oracle (the norm weight's bf16 rounding moves from the activations into W': tolerance 2^-6)."""
    g = torch.Generator().manual_seed(M * 7 + N + K)
    x = bf(torch.randn(M, K, generator=g)).cuda()
    w = bf(torch.randn(N, K, generator=g) * 0.1).cuda()
    bias = bf(torch.randn(N, generator=g)).cuda()
    res = bf(torch.randn(M, N, generator=g)).cuda()
    wp = L.PackedWeight(w)
    lin = ref_linear(x, w, bias)
    outs = []
    for rep in range(2):
        outs.append(L.gemm_stream(x, wp, pdl=bool(rep)))
        assert_close(outs[-1], ref_linear(x, w), what=f"stream plain rep {rep}")
    assert torch.equal(outs[0], outs[1]), "the stream-K reduction must be deterministic (and PDL must not change results)"
    assert_close(L.gemm_stream(x, wp, bias=bias, epi=L.EPI_GELU), torch.nn.functional.gelu(lin), what="stream gelu")
    assert_close(L.gemm_stream(x, wp, out_f32=True), ref_linear(x, w), what="stream f32 out")
    # residual in place + sums of squares of the rows written
    h = res.clone()
    ssq = torch.full((L.ssq_parts(N) * 64,), float("nan"), device="cuda")
    L.gemm_stream(x, wp, out=h, res=h, bias=bias, epi=L.EPI_RES, ssq_out=ssq)
    assert_close(h, res.float().cpu() + lin, what="stream res in place")
    got_ssq = ssq.view(-1, 64)[:, :M].sum(0).cpu()
    want_ssq = h.float().pow(2).sum(-1).cpu()
    assert torch.allclose(got_ssq, want_ssq, rtol=1e-4), "sum of squares must be that of the bf16 rows actually written"
    if N % 16 == 0:
        gate, up = w[0::2].cpu(), w[1::2].cpu()
        wantg = torch.nn.functional.silu(ref_linear(x, gate)) * ref_linear(x, up)
        assert_close(L.gemm_stream(x, wp, epi=L.EPI_SWIGLU), wantg, what="stream swiglu")
    # folded RMSNorm: next layer consumes h (K2 = N) through W2 * g
    if N % 8 == 0 and N <= 8192:
        N2 = 384
        gw = bf(torch.randn(N, generator=g) * 0.1 + 1.0).cuda()
        w2 = bf(torch.randn(N2, N, generator=g) * 0.1).cuda()
        wp2 = L.PackedWeight(w2, col_scale=gw)
        y = L.gemm_stream(h, wp2, ssq_in=ssq, ssq_in_parts=L.ssq_parts(N), norm_dim=N, eps=1e-6)
        want = ref_linear(O.rms_norm(h.cpu(), gw.cpu(), 1e-6), w2)
        assert_close(y, want, rel=2 ** -6, what="stream folded rmsnorm")
        ssq1 = torch.empty(64, device="cuda")
        L.row_ssq(h, ssq1, parts=1)
        assert torch.allclose(ssq1[:M].cpu(), want_ssq, rtol=1e-4)
        y1 = L.gemm_stream(h, wp2, ssq_in=ssq1, ssq_in_parts=1, norm_dim=N, eps=1e-6)
        assert_close(y1, want, rel=2 ** -6, what="stream folded rmsnorm (row_ssq)")
    flags = L._stream_workspace(x.device)[-(2 * L.num_sms() * 4 + 256):-256]
    assert int(flags.count_nonzero()) == 0, "every partial-tile flag must have been consumed"


def test_gemm_strided_views_and_errors(L):
    g = torch.Generator().manual_seed(11)
    big = bf(torch.randn(200, 1024, generator=g)).cuda()
    x = big[:, 256:768]  # row stride 1024, K = 512
    w = bf(torch.randn(256, 512, generator=g) * 0.1).cuda()
    outbuf = torch.zeros(200, 512, device="cuda", dtype=torch.bfloat16)
    L.gemm(x, w, out=outbuf[:, 128:384])
    assert_close(outbuf[:, 128:384], ref_linear(x, w), what="strided")
    assert outbuf[:, :128].abs().max().item() == 0 and outbuf[:, 384:].abs().max().item() == 0
    with pytest.raises(L.OmcError):
        L.gemm(x[:, :100], w[:, :100])  # K % 8 != 0
    with pytest.raises(L.OmcError):
        L.gemm(x, w, epi=L.EPI_RES)  # residual missing


# ----------------------------------------------------------------------------------------------- row ops
def test_rmsnorm(L):
    g = torch.Generator().manual_seed(1)
    for rows, C in [(7, 3200), (1025, 3584), (3, 256)]:
        x = bf(torch.randn(rows, C, generator=g) * 3).cuda()
        w = bf(torch.randn(C, generator=g) * 0.1 + 1).cuda()
        got = L.rmsnorm(x, w, 1e-6)
        want = O.rms_norm(x.cpu(), w.cpu(), 1e-6)  # bf16 oracle path: same rounding points
        assert_close(got, want, rel=2 ** -7, what="rmsnorm")
    # in place on a strided slice (the ViT QK-norm form)
    qkv = bf(torch.randn(50, 3 * 256, generator=g)).cuda()
    ref = qkv.clone()
    w = bf(torch.randn(256, generator=g) * 0.1 + 1).cuda()
    L.rmsnorm(qkv[:, 256:512], w, 1e-6, out=qkv[:, 256:512])
    assert torch.equal(qkv[:, :256], ref[:, :256]) and torch.equal(qkv[:, 512:], ref[:, 512:])
    assert_close(qkv[:, 256:512], O.rms_norm(ref[:, 256:512].cpu(), w.cpu(), 1e-6), what="rmsnorm in place")


def test_vit_embeddings(L):
    from tiny import TINY, tiny_state_dict
    sd = tiny_state_dict(0)
    cfg = O.OracleConfig(vit_hidden=TINY["vit_hidden"], image_size=TINY["image_size"])
    g = torch.Generator().manual_seed(3)
    pix = torch.randn(3, 3, 224, 224, generator=g)
    want = O.vit_embeddings(pix, sd, cfg)
    C = TINY["vit_hidden"]
    wmat = torch.zeros(C, 640)
    wmat[:, :588] = sd[O.VT + "embeddings.patch_embedding.weight"].reshape(C, 588)
    cols = L.vit_im2col(pix.cuda())
    # im2col itself is an exact gather (up to the bf16 cast of the pixels)
    un = torch.nn.functional.unfold(pix, 14, stride=14).transpose(1, 2).reshape(-1, 588)
    assert torch.equal(cols[:, :588].cpu(), bf(un)) and cols[:, 588:].abs().max().item() == 0
    patch = L.gemm(cols, bf(wmat).cuda(), bias=bf(sd[O.VT + "embeddings.patch_embedding.bias"]).cuda())
    hidden = L.vit_assemble(patch, bf(sd[O.VT + "embeddings.class_embedding"]).reshape(-1).cuda(),
                            bf(sd[O.VT + "embeddings.position_embedding"]).reshape(-1, C).cuda(), 3)
    assert_close(hidden.view(3, 257, C).flatten(0, 1), want.flatten(0, 1), rel=2 ** -6, what="vit embeddings")


@pytest.mark.parametrize("down", [1, 2, 4])
def test_select_pixel_shuffle_bit_exact(L, down):
    B, G, C = 3, 16, 64
    hidden = bf(torch.randn(B, G * G + 1, C))
    got = L.select_pixel_shuffle(hidden.cuda(), B, G, down).view(B, (G // down) ** 2, C * down * down)
    want = O.pixel_shuffle(hidden[:, 1:], down)
    assert torch.equal(got.cpu(), want)


def test_embed_lookup_and_argmax(L):
    g = torch.Generator().manual_seed(2)
    table = bf(torch.randn(1000, 256, generator=g))
    ids = torch.randint(0, 1000, (77,), generator=g)
    assert torch.equal(L.embed_lookup(ids.cuda(), table.cuda()).cpu(), table[ids])
    logits = torch.randn(5, 152064, generator=g)
    logits[2, 777] = logits[2, 150000] = 99.0  # tie -> lowest index
    logits[4, 152063] = 100.0
    got = L.argmax(logits.cuda()).cpu()
    want = torch.tensor([int(torch.argmax(r)) for r in logits])
    want[2] = 777
    assert torch.equal(got, want)


def _splice_reference(rows, feats, table, max_len):
    plans = O.splice_plan(rows, feats.shape[0], feats.shape[1], max_len)
    emb, pos, sid, off = [], [], [], [0]
    for s, plan in enumerate(plans):
        for t, (kind, idx, row) in enumerate(plan):
            emb.append(table[idx] if kind == 0 else feats[idx, row])
            pos.append(t)
            sid.append(s)
        off.append(len(emb))
    return torch.stack(emb), torch.tensor(pos), torch.tensor(sid), torch.tensor(off)


@pytest.mark.parametrize("max_len", [0, 40])
def test_splice_bit_exact(L, max_len):
    g = torch.Generator().manual_seed(4)
    C, Limg = 64, 16
    table = bf(torch.randn(500, C, generator=g))
    feats = bf(torch.randn(6, Limg, C, generator=g))
    rows = [[5, -200, 7, 8, -200, 9], [11, 12, 13], [-200], [-200, 1, 2, 3, 4, 5, 6, 7, 8, 9, -200]]
    ids = torch.tensor([t for r in rows for t in r], dtype=torch.int64)
    offs = torch.tensor([0] + list(torch.tensor([len(r) for r in rows]).cumsum(0)), dtype=torch.int32)
    cap = ids.numel() + feats.shape[0] * (Limg - 1)
    emb, pos, sid, off = L.splice(ids.cuda(), offs.cuda(), table.cuda(), feats.cuda(), -200, max_len, cap)
    remb, rpos, rsid, roff = _splice_reference(rows, feats, table, max_len if max_len else None)
    T = int(roff[-1])
    assert torch.equal(off.cpu().long(), roff)
    assert torch.equal(pos[:T].cpu().long(), rpos) and torch.equal(sid[:T].cpu().long(), rsid)
    assert torch.equal(emb[:T].cpu(), remb)


# ----------------------------------------------------------------------------------------------- attention
def _ref_attention(q, k, v, causal, scale):
    """q [S,Hq,D], k/v [S,Hkv,D] fp32 CPU -> [S,Hq,D]; fp32 softmax (eager_attention_forward / _naive_attn)."""
    S, Hq, D = q.shape
    Hkv = k.shape[1]
    k = k.repeat_interleave(Hq // Hkv, dim=1)
    v = v.repeat_interleave(Hq // Hkv, dim=1)
    w = torch.einsum("shd,thd->hst", q, k) * scale
    if causal:
        w = w.masked_fill(torch.triu(torch.ones(S, S, dtype=torch.bool), 1), float("-inf"))
    return torch.einsum("hst,thd->shd", w.softmax(-1), v)


@pytest.mark.parametrize("legacy", [False, True])
@pytest.mark.parametrize("causal,Hq,Hkv,lens", [(False, 2, 2, [257, 257]), (False, 5, 5, [1025]), (True, 14, 2, [300, 1, 64, 129]),
                                                (True, 7, 1, [1088]), (False, 1, 1, [128]), (True, 2, 1, [256, 128]),
                                                (False, 3, 3, [144, 16, 640]), (True, 4, 2, [513, 127, 384])])
def test_attention_fwd(L, causal, Hq, Hkv, lens, legacy):
    """legacy=False: tcgen05/TMEM kernel on every full 128-row query tile + mma.sync kernel on the ragged tails (the
    product path); legacy=True: mma.sync kernel on everything. Same fp32 reference, same tolerance."""
    L.attention_set_impl(legacy)
    g = torch.Generator().manual_seed(sum(lens))
    total = sum(lens)
    q = bf(torch.randn(total, Hq, 128, generator=g))
    k = bf(torch.randn(total, Hkv, 128, generator=g))
    v = bf(torch.randn(total, Hkv, 128, generator=g))
    # packed qkv rows like the product path uses
    qkv = torch.cat([q.flatten(1), k.flatten(1), v.flatten(1)], dim=1).cuda()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    out = torch.zeros(total, Hq * 128, device="cuda", dtype=torch.bfloat16)
    scale = 128 ** -0.5
    L.attention(qkv[:, :Hq * 128], qkv[:, Hq * 128:(Hq + Hkv) * 128], qkv[:, (Hq + Hkv) * 128:], out, cu, max(lens), Hq,
                Hkv, causal, scale)
    o = 0
    for n in lens:
        want = _ref_attention(q[o:o + n].float(), k[o:o + n].float(), v[o:o + n].float(), causal, scale)
        assert_close(out[o:o + n].view(n * Hq, 128), want.reshape(n * Hq, 128), rel=2 ** -6, what=f"attention len {n}")
        o += n
    L.attention_set_impl(False)


@pytest.mark.parametrize("causal,Hq,Hkv,lens", [(False, 16, 16, [1025]), (False, 4, 4, [1025, 1025, 300]), (False, 2, 2, [64, 1, 129]),
                                                 (True, 4, 2, [700, 77]), (False, 3, 3, [2000])])
def test_attention_fwd_head_dim_64(L, causal, Hq, Hkv, lens):
    """The head_dim-64 instantiation of the tcgen05 kernel (InternViT-300M: 16 heads of 64, 1025 tokens) against the fp32
    reference: full and ragged query / key tiles, several sequences, GQA + causal for completeness."""
    g = torch.Generator().manual_seed(sum(lens) + 64)
    total, D = sum(lens), 64
    q = bf(torch.randn(total, Hq, D, generator=g))
    k = bf(torch.randn(total, Hkv, D, generator=g))
    v = bf(torch.randn(total, Hkv, D, generator=g))
    qkv = torch.cat([q.flatten(1), k.flatten(1), v.flatten(1)], dim=1).cuda()
    cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), dtype=torch.int32).cuda()
    out = torch.zeros(total, Hq * D, device="cuda", dtype=torch.bfloat16)
    scale = D ** -0.5
    L.attention(qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:], out, cu, max(lens), Hq, Hkv, causal, scale,
                head_dim=D)
    o = 0
    for n in lens:
        want = _ref_attention(q[o:o + n].float(), k[o:o + n].float(), v[o:o + n].float(), causal, scale)
        assert_close(out[o:o + n].view(n * Hq, D), want.reshape(n * Hq, D), rel=2 ** -6, what=f"attention D=64 len {n}")
        o += n
    with pytest.raises(L.OmcError):
        L.attention(qkv[:, :Hq * D], qkv[:, Hq * D:(Hq + Hkv) * D], qkv[:, (Hq + Hkv) * D:], out, cu, max(lens), Hq, Hkv, causal,
                    scale, head_dim=32)


def _make_paged(B, Hkv, ctx_max, page, g):
    pages_per = (ctx_max + page - 1) // page
    n_pages = B * pages_per + 3
    perm = torch.randperm(n_pages, generator=g)[: B * pages_per].reshape(B, pages_per).to(torch.int32)
    pool = torch.zeros(n_pages, 2, Hkv, page, 128, dtype=torch.bfloat16)
    return pool, perm


@pytest.mark.parametrize("B,Hq,Hkv,page,ctxs", [(1, 28, 4, 16, [1088]), (3, 14, 2, 64, [5, 200, 333]), (2, 2, 1, 16, [17, 16])])
def test_rope_kv_store_and_paged_decode(L, B, Hq, Hkv, page, ctxs):
    g = torch.Generator().manual_seed(sum(ctxs))
    cfg = O.OracleConfig(hidden=Hq * 128, heads=Hq, kv_heads=Hkv)
    inv = O.rope_inv_freq(cfg)
    W = (Hq + 2 * Hkv) * 128
    pool, table = _make_paged(B, Hkv, max(ctxs), page, g)
    pool_d, table_d = pool.cuda(), table.cuda()
    # prefill part: ctx-1 tokens per sequence through omc_rope_kv_store
    lens = [c - 1 for c in ctxs]
    qkv = bf(torch.randn(sum(lens) + 1, W, generator=g))[: sum(lens)]
    pos = torch.cat([torch.arange(n) for n in lens]).to(torch.int32) if sum(lens) else torch.zeros(0, dtype=torch.int32)
    sid = torch.cat([torch.full((n,), i) for i, n in enumerate(lens)]).to(torch.int32) if sum(lens) else pos
    qkv_d = qkv.clone().cuda()
    if sum(lens):
        L.rope_kv_store(qkv_d, pos.cuda(), sid.cuda(), Hq, Hkv, inv.cuda(), pool_d, table_d, page)
    # oracle rotation
    def rot(x, p):  # x [n, H, 128] fp32, p [n]
        cos, sin = O.rope_cos_sin(p[None], cfg, torch.float32)
        return x * cos[0][:, None] + O.rotate_half(x) * sin[0][:, None]
    if sum(lens):
        qr = rot(qkv[:, :Hq * 128].float().view(-1, Hq, 128), pos)
        kr = rot(qkv[:, Hq * 128:(Hq + Hkv) * 128].float().view(-1, Hkv, 128), pos)
        assert_close(qkv_d[:, :Hq * 128], qr.flatten(1), rel=2 ** -7, what="rope q")
        assert_close(qkv_d[:, Hq * 128:(Hq + Hkv) * 128], kr.flatten(1), rel=2 ** -7, what="rope k")
        assert torch.equal(qkv_d[:, (Hq + Hkv) * 128:].cpu(), qkv[:, (Hq + Hkv) * 128:])
    # decode step: new token per sequence, fused RoPE + append + attention
    new = bf(torch.randn(B, W, generator=g))
    ctx_t = torch.tensor(ctxs, dtype=torch.int32)
    splits = max(2, L.decode_attn_splits(B, Hkv, max(ctxs)))
    ws = L.decode_attn_workspace(B, Hq, Hkv, splits, "cuda")
    out = torch.zeros(B, Hq * 128, device="cuda", dtype=torch.bfloat16)
    scale = 128 ** -0.5
    for rep in range(2):  # twice: the self-resetting split counters must allow a replay
        L.paged_decode_attn(new.cuda(), inv.cuda(), pool_d, table_d, page, ctx_t.cuda(), Hq, Hkv, splits, scale, out, ws)
    torch.cuda.synchronize()
    o = 0
    for b, c in enumerate(ctxs):
        n = c - 1
        kprev = qkv_d[o:o + n, Hq * 128:(Hq + Hkv) * 128].float().cpu().view(n, Hkv, 128)
        vprev = qkv[o:o + n, (Hq + Hkv) * 128:].float().view(n, Hkv, 128)
        p_new = torch.tensor([c - 1])
        qn = bf(rot(new[b:b + 1, :Hq * 128].float().view(1, Hq, 128), p_new)).float()
        kn = bf(rot(new[b:b + 1, Hq * 128:(Hq + Hkv) * 128].float().view(1, Hkv, 128), p_new)).float()
        vn = new[b:b + 1, (Hq + Hkv) * 128:].float().view(1, Hkv, 128)
        kk, vv = torch.cat([kprev, kn]), torch.cat([vprev, vn])
        kr = kk.repeat_interleave(Hq // Hkv, dim=1)
        vr = vv.repeat_interleave(Hq // Hkv, dim=1)
        w = (torch.einsum("hd,thd->ht", qn[0], kr) * scale).softmax(-1)
        want = torch.einsum("ht,thd->hd", w, vr)
        assert_close(out[b].view(Hq, 128), want, rel=2 ** -6, what=f"paged decode b={b} ctx={c}")
        # the new token's K/V landed in the right page/slot
        pg, sl = int(table[b, (c - 1) // page]), (c - 1) % page
        assert_close(pool_d[pg, 0, :, sl], kn[0], rel=2 ** -7, what="appended k")
        assert torch.equal(pool_d[pg, 1, :, sl].cpu(), bf(vn[0]))
        o += n
    # also with a single split (direct write path)
    out1 = torch.zeros_like(out)
    ws1 = L.decode_attn_workspace(B, Hq, Hkv, 1, "cuda")
    L.paged_decode_attn(new.cuda(), inv.cuda(), pool_d, table_d, page, ctx_t.cuda(), Hq, Hkv, 1, scale, out1, ws1)
    assert_close(out1, out, rel=2 ** -6, what="splits=1 vs splits>1")


# ----------------------------------------------------------------------------------------------- GEMV
@pytest.mark.parametrize("B", [1, 2, 4])
def test_gemv_variants(L, B):
    g = torch.Generator().manual_seed(20 + B)
    K, N = 3584, 1000
    x = bf(torch.randn(B, K, generator=g)).cuda()
    w = bf(torch.randn(N, K, generator=g) * 0.05).cuda()
    bias = bf(torch.randn(N, generator=g)).cuda()
    nw = bf(torch.randn(K, generator=g) * 0.1 + 1).cuda()
    assert_close(L.gemv(x, w, bias=bias), ref_linear(x, w, bias), what="gemv bias")
    xn = O.rms_norm(x.cpu(), nw.cpu(), 1e-6)
    assert_close(L.gemv(x, w, norm_w=nw, eps=1e-6, out_f32=True), ref_linear(xn, w), what="gemv norm f32")
    h = bf(torch.randn(B, N, generator=g)).cuda()
    h0 = h.clone()
    L.gemv(x, w, out=h, res=h, epi=L.EPI_RES)
    assert_close(h, h0.float().cpu() + ref_linear(x, w), what="gemv residual in place")
    I, K2 = 512, 18944
    gate = bf(torch.randn(I, 256, generator=g) * 0.1)
    up = bf(torch.randn(I, 256, generator=g) * 0.1)
    x2 = bf(torch.randn(B, 256, generator=g)).cuda()
    got = L.gemv(x2, interleave_gate_up(gate, up).cuda(), epi=L.EPI_SWIGLU)
    assert_close(got, torch.nn.functional.silu(ref_linear(x2, gate)) * ref_linear(x2, up), what="gemv swiglu")
    x3 = bf(torch.randn(B, K2, generator=g)).cuda()
    w3 = bf(torch.randn(72, K2, generator=g) * 0.02).cuda()
    assert_close(L.gemv(x3, w3), ref_linear(x3, w3), what="gemv long K")


def test_gemm_autotune_picks_a_config_and_keeps_the_bits(L):
    """Mid-sized M (one crop / one prompt): lib.gemm times the tile configurations once per shape and caches the choice.
    Every configuration accumulates along K in the same order, so the output bits must not depend on it."""
    g = torch.Generator().manual_seed(5)
    M, N, K = 1025, 3200, 640
    x = bf(torch.randn(M, K, generator=g)).cuda()
    w = bf(torch.randn(N, K, generator=g) * 0.1).cuda()
    res = bf(torch.randn(M, N, generator=g)).cuda()
    bias = bf(torch.randn(N, generator=g)).cuda()
    L._gemm_tuned.clear()
    assert L.GEMM_AUTOTUNE and M <= L.GEMM_AUTOTUNE_MAX_M
    tuned = L.gemm(x, w, bias=bias, res=res, epi=L.EPI_RES)
    assert len(L._gemm_tuned) == 1 and next(iter(L._gemm_tuned.values())) in L._GEMM_CFGS
    assert_close(tuned, res.float().cpu() + ref_linear(x, w, bias), what="autotuned gemm")
    for cfg in L._GEMM_CFGS:
        assert torch.equal(L.gemm(x, w, bias=bias, res=res, epi=L.EPI_RES, tile_cfg=cfg), tuned), hex(cfg)
    again = L.gemm(x, w, bias=bias, res=res, epi=L.EPI_RES)
    assert torch.equal(again, tuned) and len(L._gemm_tuned) == 1
    # in place (out == res, how the decoder calls it): tuning must not touch the caller's buffers
    h = res.clone()
    L._gemm_tuned.clear()
    L.gemm(x, w, out=h, bias=bias, res=h, epi=L.EPI_RES)
    assert torch.equal(h, tuned)
