"""CPU tests (no GPU) of the N > 1 host logic: slab partition + gloo world_size-2 halo exchange."""
import os
import subprocess
import sys

from conftest import ROOT
from adpres_b200.slab import slab_planes


def test_slab_partition_covers_all_planes():
    for nzz in (2, 19, 190, 418, 7):
        for n in (1, 2, 4, 8):
            if nzz < n:
                continue
            parts = [slab_planes(nzz, n, r) for r in range(n)]
            assert parts[0][0] == 0 and parts[-1][1] == nzz
            assert all(parts[r][1] == parts[r + 1][0] for r in range(n - 1))
            sizes = [b - a for a, b in parts]
            assert max(sizes) - min(sizes) <= 1


def test_gloo_two_rank_halo_and_spmv():
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", "29517", os.path.join(ROOT, "tests", "gloo_worker.py")]
    env = dict(os.environ, OMP_NUM_THREADS="1")
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=300, env=env)
    out = r.stdout + r.stderr
    assert r.returncode == 0, out[-3000:]
    assert "GLOO RANK 0 OK" in out and "GLOO RANK 1 OK" in out, out[-3000:]
