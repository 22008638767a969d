#!/usr/bin/env python
"""BASELINE.json configs[2] with the CPU oracle: IAEA-3D radial map at 4 x 4 nodes per assembly, 190 planes
(34 x 34 x 190 = 183 160 nodes), the reference's default iteration control (nin = 2, nupd = ceil(258 / 2.5) = 104).
About a minute of CPU; the result is committed as the fixture of the GPU parity test at that size (1 GPU and z-slabs).
usage: python tools/c3_oracle.py <out.json> [serc=ferc] [nin]
Round 2: the committed fixture is converged to serc = ferc = 1e-8 -- with the default 1e-5 the solution at the exit
iteration still moves by 1.2e-5 per 20 iterations (nin = 2 sweeps are far from converged), more than the 1e-5 bar on the
assembly power, so two summation orders cannot be compared there."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adpres_b200.deck import Problem
from oracle import Oracle
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
with open(os.path.join(ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh)).refine(xdiv=[2] + [4] * 8, ydiv=[4] * 8 + [2], zdiv=[10] * 19)
tol = float(sys.argv[2]) if len(sys.argv) > 2 else None
# nin: the reference default for this deck is 2.  On this mesh (5 cm x 5 cm x 2 cm nodes) the two-node iteration is then
# only marginally stable: the same solve with another grouping of the partial sums takes 1 461 ... 2 573 outers with source-
# error excursions of 1e3 ... 1e5, and one order ran into the reference's own STOP (ndmax > 1e3) -- tools/order_probe.py.
# From nin = 4 on every order converges smoothly (602 - 617 outers), so the committed fixture uses nin = 4.
nin = int(sys.argv[3]) if len(sys.argv) > 3 else None
o = Oracle(p, nout=30000, serc=tol, ferc=tol, nin=nin)
t0 = time.time()
rc, n = o.outer(0)
dt = time.time() - t0
rc2, pw = o.powdis()
res = {"what": "CPU oracle, IAEA-3D at 4 x 4 nodes per assembly and 190 planes (BASELINE configs[2]), default %ITER",
       "xdiv": [2] + [4] * 8, "ydiv": [4] * 8 + [2], "zdiv": [10] * 19, "nnod": int(p.nnod), "nin": int(nin if nin is not None else p.nin), "nupd": int(p.nupd),
       "serc": float(p.serc if tol is None else tol),
       "status": int(rc), "outers": int(n), "keff": o.state()["Ke"], "seconds": dt, "nodal_updates": o.nodal_trace(),
       "asm_power": p.asm_power(pw).tolist(),
       "power_samples": {str(i): float(pw[i]) for i in range(0, p.nnod, max(1, p.nnod // 997))}}
with open(sys.argv[1], "w") as fh:
    json.dump(res, fh)
print("done", n, o.state()["Ke"], dt)
