#!/usr/bin/env python
"""BASELINE.json configs[3] (8 energy groups, ADFs on every face; synthetic cross sections of tests/synth.py) on a 5 cm
mesh (34 x 34 x 76 = 73 264 nodes, 586 k unknowns) with the CPU oracle, nin = 10, nupd = 50 (the reference default
nin = 2 diverges on this problem before the first nodal update).  80 s of CPU; committed as the fixture of the GPU parity
test that runs the 16-lane two-node kernel at scale.   usage: python tools/c4_mid_oracle.py <out.json>"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
from synth import iaea3d_multigroup
from oracle import Oracle
cfg = dict(ng=8, xdiv=[2] + [4] * 8, ydiv=[4] * 8 + [2], zdiv=[4] * 19, nin=10, nupd=50, nac=5)
p = iaea3d_multigroup(cfg["ng"]).refine(xdiv=cfg["xdiv"], ydiv=cfg["ydiv"], zdiv=cfg["zdiv"])
o = Oracle(p, nin=cfg["nin"], nupd=cfg["nupd"], nac=cfg["nac"], nout=4000)
t0 = time.time()
rc, n = o.outer(0)
dt = time.time() - t0
rc2, pw = o.powdis()
res = dict(cfg, what="CPU oracle, synthetic 8-group IAEA-3D with ADFs (tests/synth.py) on a 5 cm mesh", nnod=int(p.nnod),
           status=int(rc), outers=int(n), keff=o.state()["Ke"], seconds=dt, nodal_updates=o.nodal_trace(),
           asm_power=p.asm_power(pw).tolist(),
           power_samples={str(i): float(pw[i]) for i in range(0, p.nnod, max(1, p.nnod // 997))})
with open(sys.argv[1], "w") as fh:
    json.dump(res, fh)
print("done", n, o.state()["Ke"], dt)
