#!/usr/bin/env python
"""BASELINE.json configs[4]: smpl/transient/LMW refined to 1 cm x 1 cm x 2 cm (110x110x100,
1.17 M nodes), rod-ejection time steps with everything device-resident (adp_xs_update,
adp_begin_time_step, outer_tr, adp_upden, adp_powtot, adp_reactivity).  Prints seconds per step.
usage: python tools/lmw_refined.py [nsteps] [nin] [nupd]"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from adpres_b200 import capi, transient
from adpres_b200.deck import Problem
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
nsteps = int(sys.argv[1]) if len(sys.argv) > 1 else 4
nin = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nupd = int(sys.argv[3]) if len(sys.argv) > 3 else 50
with open(os.path.join(ROOT, "tests", "golden", "LMW.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh))
p = p.refine(xdiv=[10] + [20] * 5, ydiv=[10] + [20] * 5, zdiv=[10] * 10)
p.nin, p.nupd, p.nac, p.nout = nin, nupd, 5, 5000
print("mesh", p.nxx, p.nyy, p.nzz, "nodes", p.nnod, "nin", nin, "nupd", nupd, flush=True)
s = capi.Solver(p)
t0 = time.perf_counter()
stamps = []
def log(msg):
    stamps.append(time.perf_counter()); print(msg, "  wall %.2f s" % (stamps[-1] - t0), flush=True)
tr = transient.rod_eject_device_glue(p, s, max_steps=nsteps, log=log, device_xs=True)
steady = stamps[0] - t0 if stamps else float("nan")
per = [(b - a) for a, b in zip(stamps[:-1], stamps[1:])]
print("steady-state phases + first step: %.2f s; later steps: %s s" % (steady, ["%.3f" % x for x in per]))
print([(r[0], r[1], round(r[2], 5), round(r[3], 6), r[4], r[5]) for r in tr])
