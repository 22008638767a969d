"""z-slab decomposition over 2 GPUs (NCCL halo planes + scalar all-reduces) against the
single-process oracle.  Needs >= 2 GPUs; run with `gpurun --gpus 2`.

(A z-refined IAEA2D -- a 2-D problem replicated axially -- is deliberately NOT used: the
reference algorithm itself is unstable on it; summing its dot products four-way instead of
serially makes the CPU oracle STOP with ndmax > 1e3, so no iteration-path parity exists.)"""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("deck", ["IAEA3Ds", "IAEA3Ds_z2"])
def test_two_rank_slabs_match_oracle(deck):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mp_worker.py"), deck]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "RANK 0/2 OK" in out and "RANK 1/2 OK" in out, out[-4000:]
