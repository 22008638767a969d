#!/usr/bin/env python
"""Is the blow-up seen in the N=2 weak-scaling bench (1 cm axial mesh) a property of the
algorithm / iteration control or of the multi-rank path?  Run the same mesh on ONE GPU."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from adpres_b200 import capi
pf = int(sys.argv[1]) if len(sys.argv) > 1 else 2
nupd = int(sys.argv[2]) if len(sys.argv) > 2 else 50
nin = int(sys.argv[3]) if len(sys.argv) > 3 else 2
p = bench.load_c2(stack=pf)
ctl = dict(bench.CTL); ctl["nupd"] = nupd; ctl["nin"] = nin
s = capi.Solver(p, **ctl)
s.matrix_setup(1); s.init_flux(); s.outer_begin(0)
nmax = int(sys.argv[4]) if len(sys.argv) > 4 else 450
for q in range(1, nmax + 1):
    ke, ser, fer = s.outer_iter(0, q)
    if q % nupd == 0:
        rc, nd, loc = s.nodal_upd(1)
        print(f"p={q:4d} ke={ke:.6f} ser={ser:.3e} fer={fer:.3e}  nodal ndmax={nd:.4e} at {loc} rc={rc}", flush=True)
    if not (ke == ke) or abs(ke) > 1e3:
        print("diverged at", q, ke); break
