"""Tensor parallelism on REAL GPUs (skipped on a single-GPU box): the model API under torchrun, one rank per GPU over NCCL +
NVLink peer memory - tools/tp_forward_check.py as a test. forward() must return full-vocabulary logits equal across ranks and
equal (cosine >= 0.999) to the unsharded model's, the data-parallel vision path must equal the replicated tower bit for bit,
and generate() at batch 8 (all-reduce fused into the GEMM epilogues as flag-in-data packets) must give the same ids on every
rank. The single-GPU emulation of the same protocols is tests/tp_stream_emulation.py / tp_mega_emulation.py."""
import json
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [2, 4])
def test_tp_forward_and_generate_on_real_gpus(world):
    if not torch.cuda.is_available() or torch.cuda.device_count() < world:
        pytest.skip(f"needs {world} GPUs")
    env = dict(os.environ, NCCL_DEBUG="WARN")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(29600 + world), os.path.join(ROOT, "tools", "tp_forward_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=420, env=env, cwd=ROOT)
    lines = [ln for ln in r.stdout.splitlines() if ln.startswith("{")]
    assert r.returncode == 0 and lines, (r.returncode, r.stdout[-2000:], r.stderr[-2000:])
    out = json.loads(lines[-1])
    assert out["ok"], out
