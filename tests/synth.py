"""Synthetic multigroup problems for the parity tests (BASELINE.json configs[3]: 8 energy groups
with assembly discontinuity factors).  The rule follows SURVEY.md section 8(d) C4: the IAEA-3D
2-group set is expanded to 4 fast + 4 thermal groups, transport cross sections grow 5 % per
group, scattering goes only to the next group, fission neutrons are born in the three fastest
groups, ADFs are 1 +- 0.05 per material / group / face from a fixed seed."""
import json
import os

import numpy as np

from conftest import GOLDEN


def iaea3d_multigroup(ng=8, seed=20261017, adf=True, zdiv=None):
    from adpres_b200.deck import Problem
    assert ng % 2 == 0 and ng >= 4
    with open(os.path.join(GOLDEN, "IAEA3Ds.spec.json")) as fh:
        d = json.load(fh)
    b = Problem.from_spec(dict(d))
    nmat, h = b.nmat, ng // 2
    sigtr = np.zeros((nmat, ng)); siga = np.zeros((nmat, ng)); nuf = np.zeros((nmat, ng)); sigf = np.zeros((nmat, ng))
    chi = np.zeros((nmat, ng)); sigs = np.zeros((nmat, ng, ng))
    for m in range(nmat):
        for g in range(ng):
            src = 0 if g < h else 1                       # fast groups copy group 1, thermal groups group 2
            sigtr[m, g] = b.xsigtr[m, src] * (1.0 + 0.05 * (g % h))
            siga[m, g] = b.xsiga[m, src]
            nuf[m, g] = b.xnuf[m, src]
            sigf[m, g] = b.xsigf[m, src]
            if g + 1 < ng:                                # down-scatter to the next group only
                sigs[m, g, g + 1] = b.xsigs[m, 0, 1] if g < h else 0.25 * b.xsiga[m, 1]
        if b.chi[m, 0] > 0:
            chi[m, :3] = (0.6, 0.3, 0.1)
    d.update(ng=ng, xsigtr=sigtr.tolist(), xsiga=siga.tolist(), xnuf=nuf.tolist(), xsigf=sigf.tolist(), chi=chi.tolist(),
             xsigs=sigs.tolist())
    if adf:
        rng = np.random.default_rng(seed)
        d["mdc"] = (1.0 + 0.05 * (2.0 * rng.random((nmat, ng, 6)) - 1.0)).tolist()
        d["adf_rot"] = []
    p = Problem.from_spec(d)
    if zdiv is not None:
        p = p.refine(zdiv=zdiv)
    return p
