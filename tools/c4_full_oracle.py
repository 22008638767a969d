#!/usr/bin/env python
"""BASELINE.json configs[3] AT SIZE with the CPU oracle: 8 energy groups + ADFs (synthetic set of tests/synth.py) on the
1 cm x 1 cm x 2 cm IAEA-3D mesh of configs[1] (170 x 170 x 190 = 4 579 000 nodes, 36.6 M unknowns, 13.7 M two-node
16 x 16 systems per nodal update).  A converged solve is out of reach for one CPU core (about 2 000 outer iterations at
~25 s each), so the fixture is a FIXED iteration count: `nout` outer iterations (nin = 10) from flat flux with a nodal
update every `nupd` -- early enough that the GPU's and the oracle's iterates have not drifted apart through round-off
(C2: |dKe| 1e-10 at p = 1..3, 1e-8 at p = 5), yet with two full SANM updates (the second one starting from non-zero dn).
Recorded: k-eff of every iteration, ndmax and location of both updates, samples of the flux and of dn.
usage: python tools/c4_full_oracle.py <out.json> [nout=10] [nupd=5]"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
from synth import iaea3d_multigroup
from oracle import Oracle
nout = int(sys.argv[2]) if len(sys.argv) > 2 else 10
nupd = int(sys.argv[3]) if len(sys.argv) > 3 else 5
cfg = dict(ng=8, xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=[10] * 19, nin=10, nupd=nupd, nac=5, nout=nout)
t0 = time.time()
p = iaea3d_multigroup(cfg["ng"]).refine(xdiv=cfg["xdiv"], ydiv=cfg["ydiv"], zdiv=cfg["zdiv"])
print("problem built", p.nnod, time.time() - t0, flush=True)
o = Oracle(p, nin=cfg["nin"], nupd=cfg["nupd"], nac=cfg["nac"], nout=nout, serc=0.0, ferc=0.0)
t1 = time.time()
rc, n = o.outer(0)
dt = time.time() - t1
ke, ser, fer = o.trace()
fdm, nod = o.times()
st = o.state()
df, dn = o.nod()                                   # (6, N, G)
rng = np.random.default_rng(4)
nodes = np.sort(rng.choice(p.nnod, 2000, replace=False))
res = dict(cfg, what="CPU oracle, synthetic 8-group IAEA-3D with ADFs (tests/synth.py) on the 1 cm x 1 cm x 2 cm mesh, fixed outer count",
           nnod=int(p.nnod), status=int(rc), outers=int(n), keff=st["Ke"], seconds=dt, cmfd_seconds=fdm, nodal_seconds=nod,
           trace_ke=ke.tolist(), trace_ser=ser.tolist(), trace_fer=fer.tolist(), nodal_updates=o.nodal_trace(),
           sample_nodes=nodes.tolist(), f0_samples=st["f0"][nodes, :].tolist(), dn_samples=dn[:, nodes, :].tolist(),
           f0_max=float(np.abs(st["f0"]).max()), dn_absmax=float(np.abs(dn).max()))
with open(sys.argv[1], "w") as fh:
    json.dump(res, fh)
print("done", n, st["Ke"], dt, fdm, nod, flush=True)
