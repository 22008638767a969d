#!/usr/bin/env python
"""Timing of the %XTAB cross-section update (adp_xs_update_xtab, k_xs_update_xtab) on the MOX part-3 core refined
radially and axially (default 8 x 8 nodes per assembly, 10 planes per axial assembly: 17 x 17 x 22 -> ~3.6 M nodes),
next to the numpy XStab_updt of the harness.  Per node the kernel writes D, sigr, nuf, sigf (G each), sigs (G x G) and
dc (6 G): 8 (4G + G^2 + 6G) = 192 B at G = 2 and reads 3 TH fields (24 B); the tables stay in L1/L2.
usage: python tools/xtab_time.py [radial_div] [axial_div] [reps]"""
import json, os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adpres_b200 import capi
from adpres_b200.deck import Problem
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
rdiv = int(sys.argv[1]) if len(sys.argv) > 1 else 8
zdiv = int(sys.argv[2]) if len(sys.argv) > 2 else 10
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 20
with open(os.path.join(ROOT, "tests", "golden", "MOX_P3_HELIOS.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh))
p = p.refine(xdiv=[rdiv // 2] + [rdiv] * (p.nx - 1), ydiv=[rdiv // 2] + [rdiv] * (p.ny - 1), zdiv=[zdiv] * p.nz)
print("mesh", p.nxx, p.nyy, p.nzz, "nodes", p.nnod, flush=True)
rng = np.random.default_rng(1)
n = p.nnod
ftem, mtem, cden = 560.0 + 700.0 * rng.random(n), np.full(n, 560.0), 0.67 + 0.08 * rng.random(n)
bpos = np.array([100.5, 0.0, 37.0, 0.0, 200.0, 200.0, 150.0, 200.0])
s = capi.Solver(p)
s.set_xtab(p); s.set_crod_map(p)
t0 = time.perf_counter()
p.update_xs(bpos, bcon=1341.99, ftem=ftem, mtem=mtem, cden=cden)
t_np = time.perf_counter() - t0
assert s.xs_update_xtab(1341.99, ftem, mtem, cden, bpos) == 0          # uploads the three fields
x = s.get_xs(); x["dc"] = s.get_dc()
print("bit-exact against numpy:", all(np.array_equal(x[k], getattr(p, k)) for k in ("D", "sigr", "nuf", "sigf", "sigs", "dc")))
s.set_th(p.th_setup())
s.set_th_state(dict(tfm=np.full((n, 13), 900.0, order="F"), heatf=np.zeros(n), ent=np.zeros(n), ftem=ftem, mtem=mtem, cden=cden))
s.xs_update_xtab(1341.99, bpos=bpos)
t0 = time.perf_counter()
for _ in range(reps):
    s.xs_update_xtab(1341.99, bpos=bpos)                                 # TH fields resident: one kernel + a flag read-back
dt = (time.perf_counter() - t0) / reps
G = p.ng
b = 8.0 * (4 * G + G * G + 6 * G) + 24.0
print("numpy XStab_updt %.2f s; device %.3f ms per update (wall, incl. launch + sync) = %.0f GB/s on %.0f B/node"
      % (t_np, 1e3 * dt, n * b / dt / 1e9, b))
