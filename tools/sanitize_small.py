#!/usr/bin/env python
"""Small runs of every round-2 kernel form for compute-sanitizer: IAEA3Ds (G = 2, SANM) 12 outers with a nodal update, with each
C / B formulation; the synthetic 8-group deck with ADFs through the quad, 16-lane and per-thread nodal kernels."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import json
from adpres_b200 import capi
from adpres_b200.deck import Problem
from synth import iaea3d_multigroup
with open(os.path.join(ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh))
for st, sp in ((6, 4), (0, 0), (2, 1), (5, 3)):
    s = capi.Solver(p, nout=12, nupd=5)
    s.set_option("st_var", st); s.set_option("spmv_var", sp)
    rc, n = s.outer(1)
    print("G=2 st_var", st, "spmv_var", sp, rc, n, s.state()["Ke"])
    s.close()
p8 = iaea3d_multigroup(8)
for form in (2, 1, 0):
    s = capi.Solver(p8, nout=7, nupd=3, nin=4)
    s.set_option("nodal_coop", form)
    rc, n = s.outer(1)
    print("G=8 nodal form", form, rc, n, s.state()["Ke"])
    s.close()
