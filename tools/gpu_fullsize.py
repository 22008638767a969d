#!/usr/bin/env python
"""Full eigenvalue solve of a refined IAEA-3D mesh on one GPU through adp_outer().
usage: python tools/gpu_fullsize.py <zdiv> [nin] [nupd]   (zdiv 10 -> C2 4.58 M nodes, 22 -> C2' 10.07 M nodes)"""
import json, os, sys, time
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from adpres_b200 import capi
from adpres_b200.deck import Problem
zdiv = int(sys.argv[1]); nin = int(sys.argv[2]) if len(sys.argv) > 2 else bench.CTL["nin"]; nupd = int(sys.argv[3]) if len(sys.argv) > 3 else bench.CTL["nupd"]
with open(os.path.join(bench.ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh)).refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[zdiv] * 19)
s = capi.Solver(p, **dict(bench.CTL, nin=nin, nupd=nupd, nout=3000))
s.enable_trace()
s.matrix_setup(1)
t0 = time.perf_counter(); rc, n = s.outer(1); dt = time.perf_counter() - t0
print(f"zdiv={zdiv} nodes={p.nnod} nin={nin} nupd={nupd}: rc={rc} outers={n} keff={s.state()['Ke']:.6f} seconds={dt:.2f}")
print("nodal updates:", [(u[0], float('%.3e' % u[1])) for u in s.trace_nodal])
