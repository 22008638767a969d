#!/usr/bin/env python
"""Formulations of the B kernel (option spmv_var, see k_spmv_dot in cmfd_kernels.cu) on BASELINE configs[1], one GPU:
isolated time and the outer iteration (graph replay, p = 51..99: no nodal update).  usage: python tools/spmv_ab.py [reps] [st_var]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from adpres_b200 import capi
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = bench.load_c2()
s = capi.Solver(p, **bench.CTL)
if len(sys.argv) > 2:
    s.set_option("st_var", int(sys.argv[2]))
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 5)
names = {0: "round-1 text, 32 regs", 1: "loads grouped, 32 regs", 2: "round-1 text, 40 regs", 3: "loads grouped, 40 regs",
         4: "loads grouped, 48 regs", 5: "loads grouped, 64 regs", 6: "round-1 text, 48 regs"}
for rep in range(reps):
    for var in range(7):
        s.set_option("spmv_var", var)
        b = s.bench_kernel(0, 20) * 1e3
        sp = s.bench_kernel(8, 20) * 1e3
        s.outer_steps(capi.MODE_FORWARD, 51, 10)
        s.timer_start()
        s.outer_steps(capi.MODE_FORWARD, 51, 49)
        ms = s.timer_stop() / 49
        print("spmv_var %d (%-24s): k_spmv_dot %6.2f us   plain spmv %6.2f us   step %.4f ms" % (var, names[var], b, sp, ms), flush=True)
