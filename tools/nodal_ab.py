#!/usr/bin/env python
"""Nodal update time on the C2 mesh for G groups: one thread per item (0), 16 lanes per surface (1), quad kernels (2).
usage: python tools/nodal_ab.py [ng ...]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from adpres_b200 import capi
from synth import iaea3d_multigroup
for ng in [int(a) for a in sys.argv[1:]] or [8]:
    p = iaea3d_multigroup(ng).refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[10] * 19)
    for coop in (0, 1, 2):
        if coop == 2 and ng < 3:
            continue
        s = capi.Solver(p, nin=10, nac=5, nupd=50, nout=3000)
        s.set_option("nodal_coop", coop)
        s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
        s.outer_steps(capi.MODE_FORWARD, 1, 3)
        ms = s.bench_kernel(7, 3)
        print("G = %d  nodal_coop = %d : nodal update %.2f ms (%.0f GB/s on 8(41G+G^2) B/node)" % (ng, coop, ms, p.nnod * 8 * (41 * ng + ng * ng) / ms / 1e6), flush=True)
        s.close()
