"""A stand-in for capi.Solver on machines without a GPU -- TEST INFRASTRUCTURE.

The device-glue drivers of the harness (thermal.DeviceGlue, transient.rod_eject_th_device, ...) talk to
capi.Solver; on the CPU they cannot run.  FakeDeviceSolver offers the same methods on top of the C oracle, the
numpy cross-section update of deck.py and oracle/th.py, keeping every array "resident" inside itself the way the
device does (TH fields, precursors, leakage, adjoint flux).  With it the Python side of the device-resident paths
-- argument plumbing, call order, which quantities come back -- is exercised on the CPU, and because the
precursor / source / reactivity arithmetic here is the C oracle's while the host-glue drivers do theirs in
numpy, the two independent restatements check each other.  It is not a fallback: nothing outside tests/ imports it."""
import numpy as np

from adpres_b200 import transient
from oracle import Oracle, th as oth


class FakeDeviceSolver:
    def __init__(self, p):
        self.p, self.o = p, Oracle(p)
        self.N, self.G = p.nnod, p.ng
        self.st = self.th = self.pline = self.af = self.sigrp = None
        self.fields = None
        self.omeg = np.zeros((p.nnod, p.ng), order="F")
        self.velo = None
        self.kin_xtab = False
        self.err = ""
        self.calls = []

    # ---- set-up calls: the data already sit in self.p
    def _note(self, name):
        self.calls.append(name)

    def set_xtab(self, p=None):
        assert (p or self.p).xtab is not None
        self._note("set_xtab")

    def set_crod_map(self, p=None):
        assert (p or self.p).crod is not None and "dsigtr" not in {k for k, v in (p or self.p).crod.items() if v is None}
        self._note("set_crod_map")

    def set_material_xs(self, p=None):
        self._note("set_material_xs")

    def set_crod(self, p=None):
        self._note("set_crod")

    def set_feedback(self, p=None):
        self._note("set_feedback")

    def set_control(self, **kw):
        self.o.set_control(**kw)

    def set_th(self, th):
        self.th = th

    def set_th_state(self, st):
        self.st = oth.initial_state(self.p, self.th)
        for k, v in st.items():
            if v is not None:
                self.st[k] = np.array(v, dtype=np.float64, order="F")

    def th_state(self):
        return {k: (None if v is None else v.copy()) for k, v in self.st.items()}

    def last_error(self):
        return self.err

    # ---- cross-section updates ("on the device": fields default to the resident TH state)
    def _xs(self, bcon, ftem, mtem, cden, bpos):
        p = self.p
        f = [a if a is not None else (None if self.st is None else self.st[k]) for a, k in ((ftem, "ftem"), (mtem, "mtem"), (cden, "cden"))]
        assert all(a is not None for a in f), "a feedback parameter is neither passed nor on the device"
        try:
            p.update_xs(bpos, bcon=bcon, ftem=f[0], mtem=f[1], cden=f[2])
        except ValueError as e:
            self.err = str(e)
            return 6 if "OUT OF THE RANGE" in self.err else 7 if "CONTROL ROD DATA" in self.err else 8
        self.o.set_xs(D=p.D, sigr=p.sigr, nuf=p.nuf, sigf=p.sigf, sigs=p.sigs, chi=p.chi, dc=p.dc, exsrc=p.exsrc)
        return 0

    def xs_update_xtab(self, bcon, ftem=None, mtem=None, cden=None, bpos=None):
        assert self.p.xtab is not None and "set_xtab" in self.calls
        return self._xs(bcon, ftem, mtem, cden, bpos)

    def xs_update_th(self, bcon, ftem=None, mtem=None, cden=None, bpos=None):
        assert self.p.xtab is None
        return self._xs(bcon, ftem, mtem, cden, bpos)

    def get_xs(self):
        p = self.p
        return dict(D=p.D.copy(), sigr=p.sigr.copy(), nuf=p.nuf.copy(), sigf=p.sigf.copy(), sigs=p.sigs.copy())

    # ---- solves
    def outer(self, popt=1):
        return self.o.outer(popt)

    def outer_ad(self, popt=1):
        return self.o.outer_ad(popt)

    def outer_th(self, maxn):
        return self.o.outer_th(maxn)

    def outer_tr(self, ht):
        return self.o.outer_tr(ht)

    def state(self):
        return self.o.state()

    def powdis(self, fixedsrc=False):
        return self.o.powdis()

    # ---- thermal hydraulics
    def th_pline(self, pow_, ppow, form=0):
        rc, npow = self.o.powdis()
        if rc:
            self.err = "TOTAL NODES POWER IS ZERO OR LESS"
            return rc
        p = self.p
        nf = self.th["node_nf"][p.ix - 1, p.iy - 1]
        if form == 0:      # th_iter, mod_th.f90:61-64
            self.pline = npow * pow_ * ppow * 0.01 / (nf * p.zdel[p.iz - 1])
        else:              # trans_calc, mod_trans.f90:430-441 (ppow = xppow, already a fraction)
            self.pline = npow * pow_ * ppow / (nf * p.zdel[p.iz - 1])
        return 0

    def th_upd(self, xpline=None, want_err=True):
        old = self.st["ftem"].copy()
        oth.th_upd(self.p, self.th, self.st, self.pline if xpline is None else xpline)
        return 0, oth.abs_e(self.st["ftem"], old)

    def th_trans(self, xpline, h):
        oth.th_trans(self.p, self.th, self.st, self.pline if xpline is None else xpline, h)
        return 0

    # ---- transient glue
    def save_adjoint(self):
        self.af = self.o.state()["f0"].copy()

    def set_kinetics(self, ibeta, lamb, velo, tbeta, sth, bth):
        self.o.set_kinetics(ibeta, lamb, velo, tbeta, sth, bth)
        self.velo, self.kin_xtab = np.asarray(velo, dtype=np.float64), False

    def set_kinetics_xtab(self, mibeta, mlamb, mvelo, tbeta, sth, bth):
        self.o.set_kinetics_xtab(mibeta, mlamb, mvelo, tbeta, sth, bth)
        self.velo, self.kin_xtab = np.asarray(mvelo, dtype=np.float64), True

    def powtot(self):
        return transient.powtot(self.p, self.o.state()["f0"])

    def ipden(self):
        self.o.ipden()

    def upden(self, ht):
        self.o.upden(ht)

    def reactivity(self, use_sigrp):
        return self.o.reactivity(self.af, self.sigrp if use_sigrp else self.p.sigr)

    def update_omeg(self, ht, bextr):
        if bextr:
            self.omeg = np.asfortranarray(np.log(self.o.state()["f0"] / self.ft) / ht)
        else:
            self.omeg = np.zeros((self.N, self.G), order="F")

    def begin_time_step(self, ht):
        p = self.p
        st = self.o.state()
        self.sigrp = p.sigr.copy(order="F")
        sigr = p.sigr.copy(order="F")
        m = p.mat - 1
        for g in range(self.G):
            v = self.velo[m, g] if self.kin_xtab else self.velo[g]
            sigr[:, g] = sigr[:, g] + 1.0 / (p.sth * v * ht) + self.omeg[:, g] / v
        self.ft, fst = st["f0"].copy(order="F"), st["fs0"].copy()
        self.o.set_xs(sigr=sigr)
        self.o.set_transient(ft=self.ft, fst=fst, omeg=self.omeg, sigrp=self.sigrp)     # c0 and L stay resident
