#!/usr/bin/env python
"""BASELINE.json configs[4] at size with the CPU oracle: smpl/transient/LMW refined to 1 cm x 1 cm x 2 cm
(110 x 110 x 100, 1.17 M nodes), steady state + adjoint + the first time steps of the rod ejection, every solve
converged to serc = ferc (default 1e-8) so that the trace does not depend on the exit iteration.
Takes tens of minutes of CPU; the result is committed as a fixture for the GPU parity test at that size.
usage: python tools/lmw_refined_oracle.py <nsteps> <out.json> [serc] [radial_div] [axial_div] [step_tol]
step_tol: serc = ferc of the time steps when it differs from the one of the steady state / adjoint (round 2: on the
1 cm mesh the t = 0 state converges to 1e-8 without trouble, only the time steps stall there)."""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adpres_b200 import transient
from adpres_b200.deck import Problem
from oracle import Oracle
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
nsteps, out = int(sys.argv[1]), sys.argv[2]
tol = float(sys.argv[3]) if len(sys.argv) > 3 else 1e-8
rdiv = int(sys.argv[4]) if len(sys.argv) > 4 else 20
zdiv = int(sys.argv[5]) if len(sys.argv) > 5 else 10
step_tol = float(sys.argv[6]) if len(sys.argv) > 6 else None
with open(os.path.join(ROOT, "tests", "golden", "LMW.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh))
p = p.refine(xdiv=[rdiv // 2] + [rdiv] * 5, ydiv=[rdiv // 2] + [rdiv] * 5, zdiv=[zdiv] * 10)
p.nin, p.nupd, p.nac, p.nout, p.serc, p.ferc, p.biter = 10, 50, 5, 20000, tol, tol, 1
print("mesh", p.nxx, p.nyy, p.nzz, "nodes", p.nnod, flush=True)
t0 = time.time()
log = []
def say(msg):
    log.append("%8.1f s  %s" % (time.time() - t0, msg)); print(log[-1], flush=True)
tr = transient.rod_eject(p, Oracle(p), max_steps=nsteps, log=say, step_tol=step_tol)
res = {"what": "CPU oracle, smpl/transient/LMW refined (xdiv %d, zdiv %d), nin=10 nupd=50 nac=5, steady state / adjoint serc=ferc=%g, "
               "time steps serc=ferc=%g, first %d time steps" % (rdiv, zdiv, tol, tol if step_tol is None else step_tol, nsteps),
       "rdiv": rdiv, "zdiv": zdiv, "nnod": int(p.nnod), "serc": tol, "step_tol": step_tol, "seconds": time.time() - t0,
       "trace": [[int(r[0]), float(r[1]), float(r[2]), float(r[3]), int(r[4]), bool(r[5])] for r in tr], "log": log}
with open(out, "w") as fh:
    json.dump(res, fh, indent=1)
print("done", time.time() - t0)
