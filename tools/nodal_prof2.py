#!/usr/bin/env python
"""One warm + one profiled nodal update at G = 2 on the C2 mesh (fused kernels) for ncu."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from adpres_b200 import capi
import bench
p = bench.load_c2()
s = capi.Solver(p, **bench.CTL)
s.set_option("graphs", 0)
s.set_option("bench_warmup", 1)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 2)
print(s.bench_kernel(7, 1))
