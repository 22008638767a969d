#!/usr/bin/env python
"""One nodal update at G groups on the C2 mesh with the kernel forms given (default 2 = quad, 1 = 16 lanes), for ncu.
usage: python tools/nodal_prof.py [ng] [form ...]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from adpres_b200 import capi
from synth import iaea3d_multigroup
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 8
p = iaea3d_multigroup(ng).refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[10] * 19)
s = capi.Solver(p, nin=2, nac=5, nupd=50, nout=3000)
s.set_option("graphs", 0)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 2)
for coop in ([int(a) for a in sys.argv[2:]] or [2, 1]):
    s.set_option("nodal_coop", coop)
    s.set_option("bench_warmup", 0)
    print(coop, s.bench_kernel(7, 1))
