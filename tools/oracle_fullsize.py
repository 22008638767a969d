#!/usr/bin/env python
"""Run the CPU oracle on a full-size refined IAEA-3D mesh (minutes to an hour of CPU) and store
the result as a committed fixture for the GPU parity test at BASELINE size.
usage: python tools/oracle_fullsize.py <zdiv> <out.json> [nin]     (zdiv 10 -> C2, 22 -> C2'; nin overrides bench.CTL)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np
import bench
from adpres_b200.deck import Problem
from oracle import Oracle

zdiv, out = int(sys.argv[1]), sys.argv[2]
with open(os.path.join(bench.ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh)).refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[zdiv] * 19)
ctl = dict(bench.CTL, nout=3000)
if len(sys.argv) > 3:
    ctl["nin"] = int(sys.argv[3])
o = Oracle(p, **ctl)
t0 = time.time()
rc, n = o.outer(0)
dt = time.time() - t0
ke, ser, fer = o.trace()
rc2, pw = o.powdis()
asm = p.asm_power(pw)
fdm, nod = o.times()
res = {"what": "CPU oracle (oracle/adpres_oracle.c), IAEA-3D refined 1cm x 1cm x %g cm, %s" % (20.0 / zdiv, json.dumps(ctl)),
       "zdiv": zdiv, "nin": ctl["nin"], "nupd": ctl["nupd"], "nac": ctl["nac"], "nnod": int(p.nnod), "status": int(rc), "outers": int(n), "keff": o.state()["Ke"],
       "seconds": dt, "cmfd_seconds": fdm, "nodal_seconds": nod,
       "trace_ke": [float(x) for x in ke], "trace_ser": [float(x) for x in ser], "trace_fer": [float(x) for x in fer],
       "nodal_updates": o.nodal_trace(), "asm_power": asm.tolist(),
       "power_samples": {str(i): float(pw[i]) for i in range(0, p.nnod, max(1, p.nnod // 997))}}
with open(out, "w") as fh:
    json.dump(res, fh)
print("done", n, o.state()["Ke"], dt)
