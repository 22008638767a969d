#!/usr/bin/env python
"""k_st (gathered loads) vs k_st_tma (cp.async.bulk + mbarrier staging) on BASELINE configs[1]: the kernel alone
(cold L2, consecutive launches stream different groups) and the whole outer iteration (graph replay, nin = 10).
usage: python tools/tma_ab.py [reps]"""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from adpres_b200 import capi
import bench
p = bench.load_c2()
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
s = capi.Solver(p, **bench.CTL)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 5)
res = {0: [], 1: []}
kern = {0: [], 1: []}
kes = {}
for rep in range(reps):
    for tma in (0, 1):
        s.set_option("st_tma", tma)
        s.outer_steps(capi.MODE_FORWARD, 51, 5)              # graph capture + warm
        s.timer_start()
        rc, ke, _, _ = s.outer_steps(capi.MODE_FORWARD, 56, 44)
        res[tma].append(s.timer_stop() / 44)
        kern[tma].append(s.bench_kernel(1, 20))
        kes[tma] = ke
for tma in (0, 1):
    print("st_tma = %d : ms/step %s | k_st alone %s us | ke %.12f" % (tma, " ".join("%.4f" % x for x in res[tma]),
          " ".join("%.2f" % (1e3 * x) for x in kern[tma]), kes[tma]), flush=True)
