#!/usr/bin/env python
"""The timed window of bench.py's value leg (BASELINE configs[1], iterations 31..50 incl. the nodal update at p = 50), step by
step: device time of every outer iteration, first as bench.py runs it (fresh context: the nodal update at p = 50 is the first
of the process), then once more on the same context.  usage: python tools/value_leg.py [opt=value ...]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from adpres_b200 import capi
import bench
import torch
p = bench.load_c2()
s = capi.Solver(p, **bench.CTL)
for a in sys.argv[1:]:
    k, v = a.split("=")
    s.set_option(k, float(v))
for rnd in range(2):
    s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
    s.outer_steps(capi.MODE_FORWARD, 1, 30)
    torch.cuda.synchronize()
    per = []
    t0 = time.perf_counter()
    for q in range(31, 51):
        s.timer_start()
        s.outer_steps(capi.MODE_FORWARD, q, 1)
        per.append(s.timer_stop())
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    print("round %d: wall %.1f ms, device per step:" % (rnd, wall * 1e3), " ".join("%.2f" % x for x in per), flush=True)
    s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
    s.outer_steps(capi.MODE_FORWARD, 1, 30)
    torch.cuda.synchronize()
    s.timer_start()
    s.outer_steps(capi.MODE_FORWARD, 31, 20)
    print("          one call of 20 steps: %.3f ms/step" % (s.timer_stop() / 20), flush=True)
