"""CPU: the C-ABI library builds, loads, and exports every symbol include/omchat_b200.h declares (no compute calls)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built():
    from omchat_b200 import build
    return build.build()


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "omchat_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(omc_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported_and_bound(built):
    from omchat_b200 import lib
    names = declared_symbols()
    assert len(names) >= 15
    cdll = ctypes.CDLL(str(built))
    for n in names:
        assert hasattr(cdll, n), f"{n} declared in the header but not exported"
        assert n in lib.SIGNATURES, f"{n} has no ctypes signature in omchat_b200/lib.py"
    assert sorted(lib.SIGNATURES) == names


def test_load_and_argument_errors_without_gpu(built):
    from omchat_b200 import lib
    l = lib.load()
    assert l.omc_version() >= 100
    # shape validation happens before any CUDA work, so it is observable on a CPU-only box
    rc = l.omc_gemm_bf16(None, 8, None, 8, None, 8, 0, 8, 8, None, None, None, 0, 0, 0, 0, None)
    assert rc == -2 and b"empty" in l.omc_last_error()
    rc = l.omc_gemm_bf16(None, 8, None, 8, None, 8, 8, 8, 12, None, None, None, 0, 0, 0, 0, None)
    assert rc == -2
    rc = l.omc_gemv_bf16(None, 8, None, 8, None, 8, 9, 8, 8, None, 1e-6, None, None, 0, 0, 0, None)
    assert rc == -2 and b"batch" in l.omc_last_error()
    rc = l.omc_rmsnorm(None, 8, None, None, 8, 4, 8192, 1e-6, None)
    assert rc == -2


def test_no_cpu_fallback():
    import torch
    from omchat_b200 import lib
    with pytest.raises(lib.OmcError):
        lib.gemm(torch.zeros(8, 8, dtype=torch.bfloat16), torch.zeros(8, 8, dtype=torch.bfloat16))
    # the variants' ops refuse host tensors just as loudly: there is no CPU path behind any of them
    x = torch.zeros(8, 64, dtype=torch.bfloat16)
    with pytest.raises(lib.OmcError):
        lib.layernorm(x, torch.ones(64, dtype=torch.bfloat16), None, 1e-6)
    with pytest.raises(lib.OmcError):
        lib.attention(x, x, x, x.clone(), torch.tensor([0, 8], dtype=torch.int32), 8, 1, 1, False, 0.125, head_dim=64)
    if not torch.cuda.is_available():
        from omchat_b200.config import OmChatQwen2MoeConfig
        from omchat_b200.model import OmChatQwen2MoeForCausalLM
        with pytest.raises(lib.OmcError):
            OmChatQwen2MoeForCausalLM(OmChatQwen2MoeConfig(num_hidden_layers=1))


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "omchat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", src, flags=re.M), f"{f} imports the oracle"


def _fake_desc(lib, n_layers=28, batch=1, hidden=3584, q=28, kv=4, inter=18944, vocab=152064, grid=148):
    import ctypes
    d = lib.DecodeDesc()
    d.n_layers, d.batch, d.hidden, d.q_heads, d.kv_heads, d.inter, d.vocab = n_layers, batch, hidden, q, kv, inter, vocab
    d.page_size, d.max_pages, d.grid, d.hist_capacity = 64, 32, grid, 16
    d.eps, d.attn_scale = 1e-6, 128 ** -0.5
    arr = (ctypes.c_void_p * n_layers)(*[0x1000 * (i + 1) for i in range(n_layers)])
    for k in ("ln1", "qkv_w", "qkv_b", "o_w", "ln2", "gate_up_w", "down_w"):
        setattr(d, k, ctypes.cast(arr, ctypes.c_void_p))
    d.rope_positions = 64 * 32
    for k in ("embed", "final_norm", "lm_head", "rope_cs", "kv_pool", "block_table", "ctx_lens", "tokens", "h", "qkv",
              "attn", "act", "logits", "workspace"):
        setattr(d, k, 0x100000)
    return d, arr


def test_decode_plan_build_is_host_only(built):
    """omc_decode_plan_build is pure CPU code: the op list of the persistent decode kernel can be checked without a GPU."""
    import ctypes
    import struct
    from omchat_b200 import lib
    l = lib.load()
    d, _keep = _fake_desc(lib)
    n = l.omc_decode_plan_bytes(28)
    buf = ctypes.create_string_buffer(n)
    assert l.omc_decode_plan_build(ctypes.byref(d), buf) == 0, l.omc_last_error()
    hdr = struct.unpack_from("16i", buf.raw, 0)
    n_ops, B, C, Hq, Hkv, G = hdr[:6]
    nslots, region_a, smem = hdr[13], hdr[14], hdr[15]
    assert (n_ops, B, C, Hq, Hkv, G) == (28 * 5 + 2, 1, 3584, 28, 4, 7)
    assert hdr[9] == 148 // 4  # key splits per (sequence, kv head)
    assert nslots >= 12 and region_a >= 18944 * 2 and smem <= 227 * 1024
    kinds = lib.mega_op_kinds(type("P", (), {"host": buf, "n_ops": n_ops}))
    assert kinds[:5] == ["qkv", "attn", "o", "gate_up", "down"] and kinds[-2:] == ["lm_head", "final"]
    assert l.omc_decode_workspace_bytes(ctypes.byref(d)) > 2 * 148 * 8 * 130 * 8
    # batch 4 still leaves a ring; batch 5 is refused; SwiGLU pair that cannot fit a ring stage is refused
    # batch 4: down_proj is cut into 3 K-chunk sub-ops (a row of 18944 elements is 3 ring slots long) so that the staged
    # activations leave room for the ring: down_0..2 with K chunks 6400 + 6400 + 6144 = 18944 and row pitch 18944
    d4, _k4 = _fake_desc(lib, batch=4)
    assert l.omc_decode_plan_build(ctypes.byref(d4), buf) == 0
    h4 = struct.unpack_from("16i", buf.raw, 0)
    assert h4[0] == 28 * 7 + 2 and h4[13] >= 8
    kinds = lib.mega_op_kinds(type("P", (), {"host": buf, "n_ops": h4[0]}))
    assert kinds[:7] == ["qkv", "attn", "o", "gate_up", "down", "down", "down"]
    op = lambda i: struct.unpack_from("6i", buf.raw, 256 + 96 * i + 56)  # N, K, ldx, ldo, kc0, ldw
    assert [op(i)[1] for i in (4, 5, 6)] == [6400, 6400, 6144] and all(op(i)[5] == 18944 for i in (4, 5, 6))
    assert op(3)[0] == 2 * 18944 and op(3)[5] == 3584
    # tune bit 3: gate_up cut as well (gate_up_0..2, down_0..2)
    d4.tune = 8
    assert l.omc_decode_plan_build(ctypes.byref(d4), buf) == 0
    h4 = struct.unpack_from("16i", buf.raw, 0)
    kinds = lib.mega_op_kinds(type("P", (), {"host": buf, "n_ops": h4[0]}))
    assert h4[0] == 28 * 9 + 2 and kinds[3:9] == ["gate_up"] * 3 + ["down"] * 3
    assert [op(i)[0] for i in (3, 4, 5)] == [12800, 12800, 12288]
    d5, _k5 = _fake_desc(lib, batch=5)
    assert l.omc_decode_plan_build(ctypes.byref(d5), buf) == -2
    dbad, _kb = _fake_desc(lib, hidden=8192)
    assert l.omc_decode_plan_build(ctypes.byref(dbad), buf) == -2


def test_header_is_plain_c_and_integration_snippet_compiles(tmp_path):
    """include/omchat_b200.h must be consumable by a C compiler (the drop-in boundary is a C ABI: plain pointers and sizes),
    and the MoE / LayerNorm / head_dim-64 calls INTEGRATION.md shows must match the declared signatures."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("no gcc")
    src = tmp_path / "hdr_check.c"
    src.write_text('''
#include <stddef.h>
#include "omchat_b200.h"
int f(void* h, void* xn, void* ln2, void* rw, void* sg, int32_t* ids, float* w, float* g, int32_t* counts, int32_t* seg,
      int32_t* cur, int32_t* te, void* xperm, void* aperm, void* yperm, int32_t* slot, void* egu, void* edn, void* sy, void* st) {
  int T = 4, C = 2048, E = 60, k = 4, I = 1408;
  int tiles = omc_moe_max_tiles(T, k, E);
  omc_moe_route(h, C, T, C, ln2, 1e-6f, xn, C, rw, sg, E, k, 0, ids, w, g, counts, st);
  omc_moe_plan_scatter(counts, E, tiles, seg, cur, te, xn, C, T, C, ids, k, xperm, C, slot, st);
  omc_gemm_bf16_grouped(xperm, C, tiles * 128, egu, C, E, 2 * I, C, te, 0, aperm, I, OMC_EPI_SWIGLU, st);
  omc_gemm_bf16_grouped(aperm, I, tiles * 128, edn, I, E, C, I, te, 0, yperm, C, OMC_EPI_NONE, st);
  omc_moe_select((const float*)h, 128, T, E, k, 0, 1, ids, w, g, counts, st);
  omc_layernorm(h, C, ln2, NULL, xn, C, T, C, 1e-6f, st);
  omc_attention_fwd_hd(h, C, h, C, h, C, xn, C, ids, 1, T, T, 16, 16, 64, 0, 0.125f, st);
  return omc_moe_combine(h, C, T, C, yperm, C, slot, w, k, sy, C, g, NULL, 1, st);
}
''')
    inc = os.path.join(ROOT, "include")
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-fsyntax-only", f"-I{inc}", str(src)],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
