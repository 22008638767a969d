"""world_size-2 CPU (gloo) check of the multi-rank host logic: slab partition, ghost-plane
exchange and all-reduced dot products reproduce the single-process SpMV / dot product."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_problem
    from adpres_b200.slab import slab_planes, exchange_halo, slab_spmv
    from oracle import Oracle
    dist.init_process_group("gloo")
    rank, world = dist.get_rank(), dist.get_world_size()
    p = load_problem("IAEA3Ds")
    o = Oracle(p)
    o.matrix_setup(1)
    a = o.matrix_dia()                       # (7, nnod, ng)
    GH, npl = 2, p.npl
    k0, k1 = slab_planes(p.nzz, world, rank)
    nown = (k1 - k0) * npl
    # plane tables exactly as adp_set_geometry builds them
    nodp = np.zeros((p.nxx + 2, p.nyy + 2), dtype=np.int64)
    nodp[p.ix[:npl], p.iy[:npl]] = np.arange(1, npl + 1)
    i, j = p.ix[:npl], p.iy[:npl]
    ypm = np.where(j == p.xstag_smin[i - 1], 0, np.arange(1, npl + 1) - nodp[i, j - 1])
    ypp = np.where(j == p.xstag_smax[i - 1], 0, nodp[i, j + 1] - np.arange(1, npl + 1))
    rng = np.random.default_rng(1)
    x = rng.standard_normal(p.nnod)
    for g in range(p.ng):
        ref = o.sp_matvec(g + 1, x)
        local = np.zeros((k1 - k0 + 2 * GH) * npl)
        local[GH * npl:GH * npl + nown] = x[k0 * npl:k1 * npl]
        exchange_halo(dist, local, npl, 1, rank, world)
        y = slab_spmv(a[:, k0 * npl:k1 * npl, g], local, ypm, ypp, npl, k1 - k0, GH)
        assert np.array_equal(y, ref[k0 * npl:k1 * npl]), "slab SpMV differs from the global one"
        part = torch.tensor([float(np.dot(y, x[k0 * npl:k1 * npl]))], dtype=torch.float64)
        dist.all_reduce(part)
        assert abs(part.item() - float(np.dot(ref, x))) < 1e-9 * abs(float(np.dot(ref, x)))
    # two-plane exchange (S3 in the nodal update)
    loc2 = np.zeros((k1 - k0 + 2 * GH) * npl)
    loc2[GH * npl:GH * npl + nown] = x[k0 * npl:k1 * npl]
    exchange_halo(dist, loc2, npl, 2, rank, world)
    lo, hi = max(0, k0 - 2), min(p.nzz, k1 + 2)
    assert np.array_equal(loc2[(lo - (k0 - GH)) * npl:(hi - (k0 - GH)) * npl], x[lo * npl:hi * npl])
    print(f"GLOO RANK {rank} OK planes [{k0},{k1})", flush=True)
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
