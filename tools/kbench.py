#!/usr/bin/env python
"""Developer micro-benchmark: per-kernel device times on the C2 mesh for a set of options.
usage: python tools/kbench.py [opt=value ...]   e.g.  fuse_st=0"""
import os, sys, json
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from adpres_b200 import capi

opts = dict(a.split("=") for a in sys.argv[1:] if "=" in a)
planes = int(opts.pop("planes", 1))
p = bench.load_c2(stack=planes)
s = capi.Solver(p, **bench.CTL)
for k, v in opts.items():
    s.set_option(k, int(v))
s.matrix_setup(1); s.init_flux(); s.outer_begin(0)
s.outer_steps(0, 1, 5)
s.timer_start(); s.outer_steps(0, 6, 40); ms = s.timer_stop()
print("opts", opts, "nin", bench.CTL["nin"], "ms/step (no nodal)", ms / 40)
s.timer_start(); s.outer_steps(0, 46, 10); ms = s.timer_stop()
print("10 steps incl. 1 nodal update: ms", ms)
names = {0: "B spmv_dot", 8: "spmv plain", 1: "C st fused", 2: "D update_xr", 3: "A update_p", 4: "P residual", 5: "F fsrc_norms",
         6: "nodal source", 7: "nodal update", 9: "matrix_setup"}
for w, n in names.items():
    reps = 3 if w in (6, 7) else 20
    print("%-14s %9.2f us" % (n, 1e3 * s.bench_kernel(w, reps)))
