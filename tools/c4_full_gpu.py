#!/usr/bin/env python
"""BASELINE configs[3] at size on the GPU against the CPU-oracle fixture tests/golden/c4_full_oracle_result.json
(tools/c4_full_oracle.py): prints the differences the test test_gpu_multigroup_adf_full_size_* bounds, and checks that the
quad nodal kernels and the one-thread-per-item kernels give bit-identical results at this size."""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from synth import iaea3d_multigroup
from adpres_b200 import capi
ref = json.load(open(os.path.join(ROOT, "tests", "golden", "c4_full_oracle_result.json")))
p = iaea3d_multigroup(ref["ng"]).refine(xdiv=ref["xdiv"], ydiv=ref["ydiv"], zdiv=ref["zdiv"])
assert p.nnod == ref["nnod"]
out = {}
for form in (2, 0):
    s = capi.Solver(p, nin=ref["nin"], nupd=ref["nupd"], nac=ref["nac"], nout=ref["outers"], serc=0.0, ferc=0.0)
    s.set_option("nodal_coop", form)
    s.enable_trace()
    t0 = time.time()
    rc, n = s.outer(0)
    dt = time.time() - t0
    st = s.state()
    out[form] = dict(rc=rc, n=n, ke=st["Ke"], f0=st["f0"], rows=list(s.trace_rows), nodal=list(s.trace_nodal), dt=dt)
    if form == 2:
        df, dn = s.nod()
        out[form]["dn"] = dn
    s.close()
a, b = out[2], out[0]
print("quad vs per-thread kernels: rc", a["rc"], b["rc"], "n", a["n"], b["n"], "seconds", a["dt"], b["dt"])
print("  bit-identical k-eff:", a["ke"] == b["ke"], " nodal trace:", a["nodal"] == b["nodal"], " flux:", np.array_equal(a["f0"], b["f0"]))
ke = np.array([r[1] for r in a["rows"]])
kr = np.array(ref["trace_ke"])
m = min(len(ke), len(kr))
d = np.abs(ke[:m] / kr[:m] - 1)
print("k-eff trace rel diff: p=1..5", d[:5], " max over all", d.max(), "at p =", int(d.argmax()) + 1, " final", d[m - 1])
print("nodal updates GPU", a["nodal"], " oracle", ref["nodal_updates"])
nodes = np.array(ref["sample_nodes"])
f0r = np.array(ref["f0_samples"])
print("flux samples: max rel diff", np.abs(a["f0"][nodes, :] / f0r - 1).max())
dnr = np.array(ref["dn_samples"])          # (6, nsample, G)
dng = a["dn"][:, nodes, :]
print("dn samples: max abs diff", np.abs(dng - dnr).max(), " max |dn|", np.abs(dnr).max(), " oracle dn_absmax", ref["dn_absmax"])
