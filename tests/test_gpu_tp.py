"""Multi-GPU tensor-parallel parity (needs ≥2 GPUs; the 1-GPU box skips it)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("world", [2, 4, 8])
def test_tp_generate_matches_oracle(world):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(29500 + world), os.path.join(ROOT, "tests", "tp_worker.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert "TP-OK" in r.stdout


@pytest.mark.parametrize("world", [2, 4, 8])
def test_single_process_group_matches_oracle(world):
    """rama_ctx_create_multi: one process, one handle, `world` devices (what a `--features gpu` caller of the reference
    can use) — decode, sampling, prefill and batched decode against the oracle."""
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "tests", "tp_single.py"), str(world)], capture_output=True, text=True,
                       timeout=900)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert f"TP1P-OK {world}" in r.stdout
