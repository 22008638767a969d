#!/usr/bin/env python
"""Formulations of the C kernel (options st_var / st_m_var, see k_st in cmfd_kernels.cu) on BASELINE configs[1], one GPU:
isolated time of k_st and of the multi-rank k_st_m (launched on one rank: no mailbox wait), and the outer iteration
(graph replay, p = 51..99: no nodal update) with each single-rank formulation.
usage: python tools/stm_ab.py [reps]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from adpres_b200 import capi
import bench
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 2
p = bench.load_c2()
s = capi.Solver(p, **bench.CTL)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 5)
names = {0: "text of round 1, 40 regs", 1: "mul.wide index", 2: "loads grouped", 3: "round-1 text, 48 regs", 4: "mul.wide, 48 regs",
         5: "loads grouped, 48 regs", 6: "loads grouped, 64 regs", 7: "round-1 text, 64 regs"}
for rep in range(reps):
    for var in range(8):
        s.set_option("st_var", var); s.set_option("st_m_var", var)
        c = s.bench_kernel(1, 20) * 1e3
        cm = s.bench_kernel(11, 20) * 1e3
        s.outer_steps(capi.MODE_FORWARD, 51, 10)
        s.timer_start()
        s.outer_steps(capi.MODE_FORWARD, 51, 49)
        ms = s.timer_stop() / 49
        print("var %d (%-24s): k_st %6.2f us   k_st_m %6.2f us   step %.4f ms" % (var, names[var], c, cm, ms), flush=True)
