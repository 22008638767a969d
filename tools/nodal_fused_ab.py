#!/usr/bin/env python
"""Nodal update time on the C2 mesh: two-kernel form (node-direction store) vs the fused per-direction kernels.
usage: python tools/nodal_fused_ab.py [ng ...]   (2 = the IAEA-3D deck itself, 4 = synthetic 4-group set)"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from adpres_b200 import capi
from synth import iaea3d_multigroup
import bench
for ng in [int(a) for a in sys.argv[1:]] or [2]:
    p = bench.load_c2() if ng == 2 else iaea3d_multigroup(ng).refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[10] * 19)
    for fused in (0, 1):
        s = capi.Solver(p, nin=10, nac=5, nupd=50, nout=3000)
        s.set_option("nodal_fused", fused)
        s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
        s.outer_steps(capi.MODE_FORWARD, 1, 3)
        ms = s.bench_kernel(7, 5)
        src = s.bench_kernel(6, 5)
        print("G = %d  nodal_fused = %d : nodal update %.3f ms (source kernel %.3f ms; %.0f GB/s on 8(41G+G^2) B/node)"
              % (ng, fused, ms, src, p.nnod * 8 * (41 * ng + ng * ng) / ms / 1e6), flush=True)
        s.close()
