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


def test_product_does_not_import_oracle():
    pkg = os.path.join(ROOT, "omchat_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith(".py"):
                src = open(os.path.join(dp, f)).read()
                assert not re.search(r"^\s*(from|import)\s+\S*oracle", src, flags=re.M), f"{f} imports the oracle"
