#!/usr/bin/env python
"""How robust is an iteration-control card against the ORDER of the global sums?  The same eigenvalue solve on one GPU with
different persistent-grid sizes (option grid_blocks: another grouping of the partial sums, nothing else changes) for several
nin: STOP code, outer count, k-eff, largest ndmax, largest source error after the tenth outer iteration.
usage: python tools/order_probe.py c3|c2|c2prime  nin[,nin...]  [grid_blocks,...]"""
import json, os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from adpres_b200 import capi
from adpres_b200.deck import Problem
import bench
cfg = sys.argv[1]
nins = [int(x) for x in sys.argv[2].split(",")]
grids = [int(x) for x in sys.argv[3].split(",")] if len(sys.argv) > 3 else [0, 600, 900, 1500]
with open(os.path.join(ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
    base = Problem.from_spec(json.load(fh))
if cfg == "c3":
    p = base.refine(xdiv=[2] + [4] * 8, ydiv=[4] * 8 + [2], zdiv=[10] * 19)
    ctl = dict(nout=30000, serc=1e-8, ferc=1e-8)
else:
    p = base.refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[22 if cfg == "c2prime" else 10] * 19)
    ctl = dict(bench.CTL, nout=3000)
print(cfg, p.nnod, "default nin", p.nin, "nupd", p.nupd, flush=True)
for nin in nins:
    for gb in grids:
        s = capi.Solver(p, **dict(ctl, nin=nin))
        if gb:
            s.set_option("grid_blocks", gb)
        s.enable_trace()
        rc, n = s.outer(1)
        nd = [x[1] for x in s.trace_nodal]
        ser = [r[2] for r in s.trace_rows[10:]]
        print("   nin=%2d grid_blocks=%4d : rc=%d outers=%5d keff=%.10f  max ndmax %.3e (%d updates)  max ser after p=10 %.3e" %
              (nin, gb, rc, n, s.state()["Ke"], max(nd) if nd else 0, len(nd), max(ser) if ser else 0), flush=True)
        s.close()
