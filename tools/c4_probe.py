#!/usr/bin/env python
"""BASELINE.json configs[3] at size: the 8-group synthetic problem with ADFs (tests/synth.py rule,
SURVEY 8(d) C4) on the C2 mesh (170x170x190, 36.6 M node-groups).  Times outer iterations and the
nodal update, then runs the eigenvalue solve.  usage: python tools/c4_probe.py [ng] [nin] [zdiv]"""
import os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from adpres_b200 import capi
from synth import iaea3d_multigroup
ng = int(sys.argv[1]) if len(sys.argv) > 1 else 8
nin = int(sys.argv[2]) if len(sys.argv) > 2 else 10
zdiv = int(sys.argv[3]) if len(sys.argv) > 3 else 10
t0 = time.time()
p = iaea3d_multigroup(ng).refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[zdiv] * 19)
print("built", p.nxx, p.nyy, p.nzz, p.nnod, "nodes x", p.ng, "groups in %.1f s" % (time.time() - t0), flush=True)
s = capi.Solver(p, nin=nin, nac=5, nupd=50, nout=3000)
s.matrix_setup(1); s.init_flux(0)
s.outer_begin(capi.MODE_FORWARD)
W, K = 5, 40
s.outer_steps(capi.MODE_FORWARD, 1, W)
s.timer_start()
rc, ke, ser, fer = s.outer_steps(capi.MODE_FORWARD, W + 1, K)
ms = s.timer_stop() / K
rows = p.nnod * p.ng
print("outer iteration: %.3f ms  -> %.3e unknowns/s/outer ; Ke %.6f after %d steps (rc %d)" % (ms, rows / ms * 1e3, ke, W + K, rc), flush=True)
nodal_ms = s.bench_kernel(7, 3)
print("nodal update (G=%d): %.3f ms = %.1f GB/s on the 8(41G+G^2) B/node basis" % (ng, nodal_ms, p.nnod * 8 * (41 * ng + ng * ng) / nodal_ms / 1e6), flush=True)
s2 = capi.Solver(p, nin=nin, nac=5, nupd=50, nout=3000)
t0 = time.perf_counter()
rc, n = s2.outer(0)
dt = time.perf_counter() - t0
print("solve: rc %d, %d outers, k-eff %.8f, %.2f s, ndmax %.3e" % (rc, n, s2.state()["Ke"], dt, s2.ndmax), flush=True)
