"""Thermal-hydraulic feedback and critical-boron search drivers for the harness (reference:
src/mod_th.f90 th_iter :11-91, cbsearch :752-837, cbsearcht :840-959).

In the drop-in deployment these loops stay Fortran and call `outer`, `outer_th`, `PowDis`,
`XS_updt`, `th_upd`.  This module restates the *callers* so that the reference's `%THER` /
`%CBCS` decks (smpl/static/NEACRP) run end to end against either back end:

    HostGlue(p, solver, th_module)   solver = oracle.Oracle (or capi.Solver): cross sections are
                                     updated in numpy (deck.Problem.update_xs) and uploaded, PowDis
                                     comes back to the host, th_upd is `th_module.th_upd` (the tests
                                     pass oracle.th)
    DeviceGlue(p, solver)            solver = capi.Solver: adp_xs_update_th (adp_xs_update_xtab for
                                     %XTAB decks), adp_outer_th, adp_th_pline, adp_th_upd -- per TH
                                     iteration one boron concentration goes up and four scalars come back

Both expose xs_update(bcon), outer(), outer_th(nth), th_step() -> th_err, and state() with Ke,
ser, fer; the drivers below are written once against that interface.
"""
from __future__ import annotations

import numpy as np


class StopError(RuntimeError):
    """one of the reference's STOP conditions"""


def _card(p, key):
    t = (p.fbk or {}).get(key)
    return None if t is None else t["val"]


def _outer_th(p, solver, nth):
    """CALL outer_th(nth).  Without a %ITER card outer_th itself sets nupd = int(nth/2), for good
    (mod_cmfd.f90:743) -- otherwise the default nupd would leave the TH iterations without a single
    nodal update."""
    if not p.biter and p.nupd != p.nth // 2:
        p.nupd = p.nth // 2
        solver.set_control(nupd=p.nupd)
    return solver.outer_th(nth)


def _initial_fields(p):
    """ftem, mtem, cden before the first TH solve: the value of the %FTEM / %MTEM / %CDEN card everywhere
    (mod_io.f90:2697-2699 etc.); %XTAB decks have no such cards and inp_ther sets 900., 500., 0.711 (:3103-3105)."""
    n = p.nnod
    if getattr(p, "xtab", None) is not None:
        d = p.xtab_defaults()
        return np.full(n, d["ftem"]), np.full(n, d["mtem"]), np.full(n, d["cden"])
    return (np.full(n, _card(p, "ftem") if _card(p, "ftem") is not None else 900.0),
            np.full(n, _card(p, "mtem") if _card(p, "mtem") is not None else 560.0),
            np.full(n, _card(p, "cden") if _card(p, "cden") is not None else 0.75))


class HostGlue:
    def __init__(self, p, solver, th_module):
        self.p, self.s, self.thm = p, solver, th_module
        self.th = p.th_setup() if p.ther is not None else None
        n = p.nnod
        # inp_ftem / inp_mtem / inp_cden: the card's value everywhere (mod_io.f90:2697-2699 etc.)
        self.ftem, self.mtem, self.cden = _initial_fields(p)
        self.bpos = None if p.crod is None else p.crod["bpos"].astype(np.float64)
        if self.th is not None:
            self.st = th_module.initial_state(p, self.th)
            self.st.update(ftem=self.ftem, mtem=self.mtem, cden=self.cden)

    def xs_update(self, bcon):
        p = self.p
        p.update_xs(self.bpos, bcon=bcon, ftem=self.ftem, mtem=self.mtem, cden=self.cden)
        self.s.set_xs(D=p.D, sigr=p.sigr, nuf=p.nuf, sigf=p.sigf, sigs=p.sigs, chi=p.chi, dc=p.dc, exsrc=p.exsrc)

    def outer(self):
        return self.s.outer(0)

    def outer_th(self, nth):
        return _outer_th(self.p, self.s, nth)

    def th_step(self):
        p, th = self.p, self.th
        rc, npow = self.s.powdis()
        if rc > 0:
            raise StopError("TOTAL NODES POWER IS ZERO OR LESS")
        pline = self.thm.pline_static(p, th, npow)
        old = self.st["ftem"].copy()
        self.thm.th_upd(p, th, self.st, pline)
        self.ftem, self.mtem, self.cden = self.st["ftem"], self.st["mtem"], self.st["cden"]
        return self.thm.abs_e(self.st["ftem"], old)

    def state(self):
        return self.s.state()

    def th_fields(self):
        return dict(ftem=self.ftem.copy(), mtem=self.mtem.copy(), cden=self.cden.copy())


class DeviceGlue:
    def __init__(self, p, solver):
        self.p, self.s = p, solver
        self.th = p.th_setup() if p.ther is not None else None
        n = p.nnod
        self.bpos = None if p.crod is None else p.crod["bpos"].astype(np.float64)
        self.xtab = getattr(p, "xtab", None) is not None
        if self.xtab:                       # %XTAB deck: branch tables, the rodded sets are part of them
            solver.set_xtab(p)
            if p.crod is not None:
                solver.set_crod_map(p)
        else:
            solver.set_material_xs(p)
            if p.crod is not None:
                solver.set_crod(p)
            solver.set_feedback(p)
        ftem, mtem, cden = _initial_fields(p)
        if self.th is not None:
            solver.set_th(self.th)
            solver.set_th_state(dict(tfm=np.full((n, self.th["nt"] + 1), 900.0, order="F"), heatf=np.zeros(n), ent=np.zeros(n),
                                     ftem=ftem, mtem=mtem, cden=cden))
            self._first = None
        else:
            self._first = (ftem, mtem, cden)          # no TH state on the device: pass the card values

    def xs_update(self, bcon):
        fields = self._first if self._first is not None else ()
        update = self.s.xs_update_xtab if self.xtab else self.s.xs_update_th
        if update(bcon, *fields, bpos=self.bpos) > 0:         # the STOPs of brInterp / crod_tab_updt / Dsigr_updt / check_xs
            raise StopError(self.s.last_error())

    def outer(self):
        return self.s.outer(0)

    def outer_th(self, nth):
        return _outer_th(self.p, self.s, nth)

    def th_step(self):
        rc = self.s.th_pline(self.th["pow"], self.th["ppow"], form=0)
        if rc > 0:
            raise StopError(self.s.last_error())
        rc, err = self.s.th_upd(None)
        if rc > 0:
            raise StopError(self.s.last_error())
        return err

    def state(self):
        return self.s.state()

    def th_fields(self):
        st = self.s.th_state()
        return dict(ftem=st["ftem"], mtem=st["mtem"], cden=st["cden"])


def th_iter(g, bcon, ind=None, log=None):
    """th_iter (mod_th.f90:11-91).  Returns (th_err, iterations)."""
    p = g.p
    mx_iter = p.th_niter if ind is not None else 2
    th_err, l = 1.0, 0
    for l in range(1, mx_iter + 1):
        g.xs_update(bcon)
        g.outer_th(p.nth)
        th_err = g.th_step()
        st = g.state()
        if log:
            log(f"   th_iter {l}: k-eff {st['Ke']:.6f} th_err {th_err:.5e} ser {st['ser']:.3e} fer {st['fer']:.3e}")
        if th_err < 0.01 and st["fer"] < p.ferc and st["ser"] < p.serc and ind is not None:
            break
    else:
        if ind is not None:
            raise StopError("MAXIMUM TH ITERATION REACHED.")
    return th_err, l


def _secant(g, rbcon, evaluate, nmax, tol_ser, tol_fer, log):
    """the search loop shared by cbsearch (:766-816) and cbsearcht (:862-911)"""
    rows = []
    bcon = rbcon
    ke = evaluate(bcon)
    rows.append((1, bcon, ke))
    bc1, ke1 = bcon, ke
    if nmax == 30 and bcon < 1.0e-5:                       # cbsearcht only (:869-873)
        bcon = 500.0
    else:
        bcon = bcon + (ke - 1.0) * bcon
    ke = evaluate(bcon)
    rows.append((2, bcon, ke))
    bc2, ke2 = bcon, ke
    n = 3
    while True:
        bcon = bc2 + (1.0 - ke2) / (ke1 - ke2) * (bc1 - bc2)
        ke = evaluate(bcon)
        bc1, bc2, ke1, ke2 = bc2, bcon, ke2, ke
        st = g.state()
        rows.append((n, bcon, ke))
        if log:
            log(f"{n:3d} {bcon:10.2f} {ke:14.5f} {st['ser']:14.5e} {st['fer']:13.5e}")
        if abs(ke - 1.0) < 1.0e-5 and st["ser"] < tol_ser and st["fer"] < tol_fer:
            break
        n += 1
        if bcon > 3000.0:
            raise StopError("CRITICAL BORON CONCENTRATION EXCEEDS THE LIMIT(3000 ppm)")
        if bcon < 0.0:
            raise StopError("CRITICAL BORON CONCENTRATION IS NOT FOUND (LESS THAN ZERO)")
        if n == nmax:
            raise StopError("MAXIMUM ITERATION FOR CRITICAL BORON SEARCH IS REACHING MAXIMUM")
    return bcon, rows


def cbsearch(g, log=None):
    """cbsearch (mod_th.f90:752-837): critical boron concentration without TH feedback."""
    def evaluate(bcon):
        g.xs_update(bcon)
        rc, n = g.outer()
        if rc > 0:
            raise StopError("outer iteration stopped with code %d" % rc)
        return g.state()["Ke"]
    return _secant(g, g.p.fbk["bcon"]["ref"], evaluate, 20, 1.0e-5, 1.0e-5, log)


def cbsearcht(g, log=None):
    """cbsearcht (mod_th.f90:840-959): critical boron concentration with TH feedback (two TH
    iterations per boron guess)."""
    def evaluate(bcon):
        th_iter(g, bcon, ind=None, log=None)
        return g.state()["Ke"]
    # rbcon: the %CBCS reference; %XTAB decks have no %CBCS card and leave it at its initial 0
    rbcon = g.p.fbk["bcon"]["ref"] if (g.p.fbk or {}).get("bcon") is not None else 0.0
    return _secant(g, rbcon, evaluate, 30, g.p.serc, g.p.ferc, log)
