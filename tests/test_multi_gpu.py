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


WORLD = int(os.environ.get("ADP_TEST_WORLD", "2"))      # 2 by default; 4 / 8 with `gpurun --gpus N`


@pytest.mark.parametrize("deck", ["IAEA3Ds", "IAEA3Ds_z2", "LMW_tr", "NEACRP_th", "NEACRP_cb", "MOX_xtab", "C3_fixture", "SYNTH8_adf", "LMW_refined"])
@pytest.mark.parametrize("mode", ["peer", "nccl"])
def test_slabs_match_oracle(deck, mode):
    """mode peer: halo planes pushed by the kernels over NVLink peer memory + mailbox all-reduce;
    mode nccl: the ncclSend/Recv + ncclAllReduce fallback path (ADP_NO_PEER=1)."""
    if _ngpu() < WORLD:
        pytest.skip(f"needs {WORLD} GPUs")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(WORLD), "--master-addr",
           "127.0.0.1", "--master-port", "29541", os.path.join(ROOT, "tests", "mp_worker.py"), deck]
    env = dict(os.environ)
    if mode == "nccl":
        env["ADP_NO_PEER"] = "1"
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-4000:]
    for q in range(WORLD):
        assert f"RANK {q}/{WORLD} OK" in out, out[-4000:]
