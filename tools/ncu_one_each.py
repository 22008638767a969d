#!/usr/bin/env python
"""Runs each hot-path kernel a few times on the C2 mesh (for `ncu --set full` captures)."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from adpres_b200 import capi
p = bench.load_c2()
s = capi.Solver(p, **bench.CTL)
s.matrix_setup(1); s.init_flux(); s.outer_begin(0)
s.outer_steps(0, 1, 3)
import torch
for w in (0, 8, 1, 2, 3, 4, 5, 6, 7, 9):
    s.bench_kernel(w, 1)            # warm (lazy module load etc.)
s.set_option("bench_warmup", 0)
torch.cuda.cudart().cudaProfilerStart()
for w in (0, 8, 1, 2, 3, 4, 5, 7, 9):
    s.bench_kernel(w, 1)
torch.cuda.cudart().cudaProfilerStop()
print("done")
