"""z-slab decomposition over 2 GPUs (NCCL halo planes + scalar all-reduces) against the
single-process oracle.  Needs >= 2 GPUs; run with `gpurun --gpus 2`."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count() if torch.cuda.is_available() else 0


@pytest.mark.parametrize("deck", ["IAEA3Ds", "IAEA2D"])
def test_two_rank_slabs_match_oracle(deck):
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29541", os.path.join(ROOT, "tests", "mp_worker.py"), deck]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    assert "RANK 0/2 OK" in out and "RANK 1/2 OK" in out, out[-4000:]
