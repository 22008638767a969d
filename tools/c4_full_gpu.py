#!/usr/bin/env python
"""BASELINE configs[3] at size on the GPU against the CPU-oracle fixture tests/golden/c4_full_oracle_result.json
(tools/c4_full_oracle.py: n1 outers from flat flux, one SANM nodal update, n2 outers): prints the differences the test
test_gpu_multigroup_adf_full_size_against_cpu_oracle_fixture bounds, for the quad nodal kernels (the default for G >= 5)
and the one-thread-per-item kernels, and checks that the two give bit-identical results at this size."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from synth import iaea3d_multigroup
from adpres_b200 import capi
ref = json.load(open(os.path.join(ROOT, "tests", "golden", "c4_full_oracle_result.json")))
p = iaea3d_multigroup(ref["ng"]).refine(xdiv=ref["xdiv"], ydiv=ref["ydiv"], zdiv=ref["zdiv"])
assert p.nnod == ref["nnod"]
nodes = np.array(ref["sample_nodes"])
out = {}
for form in (2, 0):
    ctl = dict(nin=ref["nin"], nupd=ref["nupd"], nac=ref["nac"], serc=0.0, ferc=0.0)
    s = capi.Solver(p, nout=ref["n1"], **ctl)
    s.set_option("nodal_coop", form)
    s.enable_trace()
    t0 = time.time()
    rc1, m1 = s.outer(0)
    rcn, nd, loc = s.nodal_upd(1)
    t1 = time.time()
    df, dn = s.nod()
    s.set_control(nout=ref["n2"], **ctl)
    rc2, m2 = s.outer(0)
    st = s.state()
    out[form] = dict(rc=[rc1, rcn, rc2], m=[m1, m2], nd=nd, loc=loc, ke=[r[1] for r in s.trace_rows], f0=st["f0"], dn=dn[:, nodes, :].copy(),
                     dnall=dn if form == 2 else None, dt=t1 - t0)
    s.close()
a, b = out[2], out[0]
print("status", a["rc"], "oracle", ref["status"], " outers", a["m"], "oracle", ref["outers"], " seconds incl. update", a["dt"], b["dt"])
print("quad vs per-thread kernels bit-identical: k-eff", a["ke"] == b["ke"], " ndmax/loc", (a["nd"], a["loc"]) == (b["nd"], b["loc"]),
      " dn samples", np.array_equal(a["dn"], b["dn"]), " flux", np.array_equal(a["f0"], b["f0"]))
ke, kr = np.array(a["ke"]), np.array(ref["trace_ke_first"] + ref["trace_ke"])
print("k-eff trace GPU", ke, "\n      oracle", kr, "\n      rel diff", np.abs(ke / kr - 1))
print("ndmax GPU %.15g at %s   oracle %.15g at %s   rel diff %.3e" % (a["nd"], a["loc"], ref["ndmax"], ref["ndloc"], abs(a["nd"] / ref["ndmax"] - 1)))
f0r = np.array(ref["f0_samples"])
print("flux samples: max rel diff", np.abs(a["f0"][nodes, :] / f0r - 1).max())
dnr = np.array(ref["dn_samples"])          # (6, nsample, G)
print("dn samples: max abs diff", np.abs(a["dn"] - dnr).max(), " max |dn| in samples", np.abs(dnr).max(), " max |dn| GPU overall",
      np.abs(a["dnall"]).max(), " oracle", ref["dn_absmax"])
