"""Rod-ejection transient driver for the harness (reference: src/mod_trans.f90).

In the drop-in deployment `rod_eject` / `trans_calc` stay Fortran and call `outer`, `outer_ad`
and `outer_tr` of the replaced `mod_cmfd`.  This module restates that *caller* (mode RODEJECT,
no TH feedback, %XSEC decks) so that the reference's transient decks can be run end to end
against any object exposing the hot-path entry points -- the CUDA library
(`adpres_b200.capi.Solver`) or the CPU oracle (`oracle.Oracle`):

    set_xs(**arrays)  set_kinetics(...)  set_transient(...)
    outer(popt)  outer_ad(popt)  outer_tr(ht)  state()  nod()

The glue between the calls (control-rod motion, `iPden`, `uPden`, `PowTot`, `reactivity` with
`Lxyz`) is numpy, identical for both back ends, so a trace comparison isolates the hot path.
"""
from __future__ import annotations

import numpy as np


def lxyz_total(p, f0, df, dn):
    """L(n,g) = L1 + L2 + L3 of Lxyz (mod_nodal.f90:901-1005), vectorised over all nodes.
    df, dn: (6, nnod, ng) as adp_get_nod returns them."""
    N, G, npl = p.nnod, p.ng, p.npl
    i, j, k = p.ix, p.iy, p.iz
    first_x = i == p.ystag_smin[j - 1]
    last_x = i == p.ystag_smax[j - 1]
    first_y = j == p.xstag_smin[i - 1]
    last_y = j == p.xstag_smax[i - 1]
    first_z, last_z = k == 1, k == p.nzz
    # neighbour node numbers (0-based); y neighbours through the plane table
    pos = np.zeros((p.nxx + 2, p.nyy + 2), dtype=np.int64)
    pos[p.ix[:npl], p.iy[:npl]] = np.arange(npl)
    base = (k - 1).astype(np.int64) * npl
    n = np.arange(N)
    xp, xm = np.where(last_x, n, n + 1), np.where(first_x, n, n - 1)
    yp = np.where(last_y, n, base + pos[i, np.minimum(j + 1, p.nyy + 1)])
    ym = np.where(first_y, n, base + pos[i, j - 1])
    zp, zm = np.where(last_z, n, n + npl), np.where(first_z, n, n - npl)
    bc = p.bc  # xeast, xwest, ynorth, ysouth, zbott, ztop
    L = np.zeros((N, G), order="F")
    for g in range(G):
        f = f0[:, g]

        def lead(face_p, face_m, nb_p, nb_m, last, first, bcp, bcm, h):
            dfp, dnp_, dfm, dnm = df[face_p, :, g], dn[face_p, :, g], df[face_m, :, g], dn[face_m, :, g]
            jp_int = -dfp * (f[nb_p] - f) - dnp_ * (f[nb_p] + f)
            jp_bnd = np.zeros(N) if bcp == 2 else dfp * f - dnp_ * f
            jm_int = -dfm * (f - f[nb_m]) - dnm * (f + f[nb_m])
            jm_bnd = np.zeros(N) if bcm == 2 else -dfm * f - dnm * f
            jp = np.where(last, jp_bnd, jp_int)
            jm = np.where(first, jm_bnd, jm_int)
            return (jp - jm) / h
        L1 = lead(0, 1, xp, xm, last_x, first_x, bc[0], bc[1], p.xdel[i - 1])
        L2 = lead(2, 3, yp, ym, last_y, first_y, bc[2], bc[3], p.ydel[j - 1])
        L3 = lead(4, 5, zp, zm, last_z, first_z, bc[5], bc[4], p.zdel[k - 1])
        L[:, g] = L1 + L2 + L3
    return L


def powtot(p, f0):
    """PowTot (mod_trans.f90:523-557)."""
    pw = np.zeros(p.nnod)
    for g in range(p.ng):
        pw = pw + np.maximum(f0[:, g] * p.sigf[:, g] * p.vdel, 0.0)
    return float(pw.sum())


def reactivity(p, af, sigrp, f0, fs0, L):
    """reactivity (mod_trans.f90:648-688) with L already evaluated."""
    src = rem = lea = fde = 0.0
    chi_n = p.chi[p.mat - 1, :]
    for g in range(p.ng):
        scg = np.zeros(p.nnod)
        for h in range(p.ng):
            if h != g:
                scg = scg + p.sigs[:, h, g] * f0[:, h]
        src += float((af[:, g] * (scg + chi_n[:, g] * fs0) * p.vdel).sum())
        rem += float((af[:, g] * sigrp[:, g] * f0[:, g] * p.vdel).sum())
        lea += float((af[:, g] * L[:, g] * p.vdel).sum())
        fde += float((af[:, g] * chi_n[:, g] * fs0 * p.vdel).sum())
    return (src - lea - rem) / fde


def _leakage(p, solver, f0):
    """L(n,g) for `reactivity`: on the device when the back end offers it (adp_lxyz_total), else
    from nod%df/dn copied back, as the reference does."""
    if hasattr(solver, "lxyz_total"):
        return solver.lxyz_total()
    df, dn = solver.nod()
    return lxyz_total(p, f0, df, dn)


def _push_xs(solver, p, **override):
    kw = dict(D=p.D, sigr=p.sigr, nuf=p.nuf, sigf=p.sigf, sigs=p.sigs, chi=p.chi, dc=p.dc, exsrc=p.exsrc)
    kw.update(override)
    solver.set_xs(**kw)


def rod_eject_device_glue(p, solver, max_steps=None, log=None, device_xs=False, step_tol=None):
    """The same transient with the time-step glue on the device (adp_save_adjoint, adp_ipden,
    adp_begin_time_step, adp_upden, adp_powtot, adp_reactivity): per step only the new cross
    sections go up and three scalars come back.  With device_xs the cross-section update itself
    (base_updt + crod_updt + Dsigr_updt) runs on the device too (adp_xs_update): per step only the
    bank positions go up.  Same return value as rod_eject()."""
    if device_xs:
        def _push(solver_, p_, bpos_):
            solver_.set_material_xs(p_)
            solver_.set_crod(p_)
            solver_.xs_update(bpos_)
    else:
        def _push(solver_, p_, bpos_):
            p_.update_xs(bpos_)
            _push_xs(solver_, p_)
    e, c = p.ejct, p.crod
    ibeta, lamb, velo = e["ibeta"], e["lamb"], e["velo"]
    bpos = c["bpos"].astype(np.float64).copy()
    fbpos, tmove, bspeed = e["fbpos"], e["tmove"], e["bspeed"]
    mdir = np.where(np.abs(fbpos - bpos) < 1e-5, 0, np.where(fbpos - bpos > 1e-5, 2, 1))
    _push(solver, p, bpos)
    rc, n = solver.outer(0)
    assert rc == 0, rc
    ke = solver.state()["Ke"]
    if abs(ke - 1.0) > 1e-5:                                  # KNE1
        for it in range(10):
            p.xnuf = p.xnuf / ke
            c["dnuf"] = c["dnuf"] / ke
            _push(solver, p, bpos)
            rc, n = solver.outer(0)
            ke = solver.state()["Ke"]
            if abs(ke - 1.0) < 1e-5:
                break
    solver.outer_ad(0)
    solver.save_adjoint()
    solver.outer(0)
    tbeta = np.full(p.nmat, 0.0)
    for jf in range(6):
        tbeta = tbeta + ibeta[jf]
    ctbeta = tbeta[0]
    solver.set_kinetics(ibeta, lamb, velo, tbeta, p.sth, p.bth)
    if step_tol is not None:            # harness option: the time steps use their own serc = ferc (fixtures)
        solver.set_control(serc=step_tol, ferc=step_tol)
    solver.ipden()
    tpow1 = solver.powtot()
    rho = solver.reactivity(0)
    trace = [(0, 0.0, rho / ctbeta, 1.0, 0, False)]
    steps = [(i, e["tstep1"], i * e["tstep1"]) for i in range(1, int(round(e["tdiv"] / e["tstep1"])) + 1)]
    steps += [(i, e["tstep2"], e["tdiv"] + i * e["tstep2"]) for i in range(1, int(round((e["ttot"] - e["tdiv"]) / e["tstep2"])) + 1)]
    for step, (_, ht, t2) in enumerate(steps, start=1):
        if max_steps is not None and step > max_steps:
            break
        for b in range(c["nb"]):
            if mdir[b] == 1 and t2 - tmove[b] > 1e-5 and fbpos[b] - bpos[b] < 1e-5:
                bpos[b] = max(bpos[b] - ht * bspeed[b], fbpos[b])
            elif mdir[b] == 2 and t2 - tmove[b] > 1e-5 and fbpos[b] - bpos[b] > 1e-5:
                bpos[b] = min(bpos[b] + ht * bspeed[b], fbpos[b])
        _push(solver, p, bpos)              # XS_updt result; the time terms are added on the device
        solver.update_omeg(ht, p.bextr and step > 1)          # mod_trans.f90:127-135,150-154
        solver.begin_time_step(ht)
        rc, maxi, n = solver.outer_tr(ht)
        assert rc == 0, rc
        solver.upden(ht)
        tpow2 = solver.powtot()
        rho = solver.reactivity(1)
        trace.append((step, t2, rho / ctbeta, tpow2 / tpow1, n, maxi))
        if log:
            log(f"{step:4d} {t2:10.3f} {rho / ctbeta:10.4f} {tpow2 / tpow1:15.4E}  outers {n}")
    return trace


def rod_eject(p, solver, max_steps=None, log=None, step_tol=None):
    """rod_eject (mod_trans.f90:17-160) + trans_calc (:332-479), thc = 0.  Returns a list of
    (step, t, reactivity [$], relative power, outer iterations, maxi)."""
    e, c = p.ejct, p.crod
    ibeta, lamb, velo = e["ibeta"], e["lamb"], e["velo"]
    bpos = c["bpos"].astype(np.float64).copy()
    fbpos, tmove, bspeed = e["fbpos"], e["tmove"], e["bspeed"]
    mdir = np.where(np.abs(fbpos - bpos) < 1e-5, 0, np.where(fbpos - bpos > 1e-5, 2, 1))
    say = log or (lambda *a: None)

    p.update_xs(bpos)
    _push_xs(solver, p)
    rc, n = solver.outer(0)
    assert rc == 0, rc
    ke = solver.state()["Ke"]
    say(f"steady state: {n} outers, k-eff {ke:.6f}")
    # KNE1 (mod_trans.f90:483-518): force k-eff to 1 by scaling nu*sigf
    if abs(ke - 1.0) > 1e-5:
        for it in range(10):
            p.xnuf = p.xnuf / ke
            c["dnuf"] = c["dnuf"] / ke
            p.update_xs(bpos)
            _push_xs(solver, p)
            rc, n = solver.outer(0)
            ke = solver.state()["Ke"]
            say(f"KNE1 pass {it + 1}: {n} outers, k-eff {ke:.6f}")
            if abs(ke - 1.0) < 1e-5:
                break
    rc, n = solver.outer_ad(0)
    af = solver.state()["f0"].copy()
    say(f"adjoint: {n} outers")
    rc, n = solver.outer(0)
    st = solver.state()
    f0, fs0 = st["f0"], st["fs0"]
    say(f"forward again: {n} outers, k-eff {st['Ke']:.6f}")
    c0 = np.asfortranarray((ibeta / lamb)[None, :] * fs0[:, None])        # iPden
    tpow1 = powtot(p, f0)
    tbeta = np.full(p.nmat, 0.0)
    for jf in range(6):
        tbeta = tbeta + ibeta[jf]
    ctbeta = tbeta[0]
    L = _leakage(p, solver, f0)
    rho = reactivity(p, af, p.sigr, f0, fs0, L)
    trace = [(0, 0.0, rho / ctbeta, 1.0, 0, False)]
    solver.set_kinetics(ibeta, lamb, velo, tbeta, p.sth, p.bth)
    if step_tol is not None:            # harness option: the time steps use their own serc = ferc (fixtures)
        solver.set_control(serc=step_tol, ferc=step_tol)

    steps = [(i, e["tstep1"], i * e["tstep1"]) for i in range(1, int(round(e["tdiv"] / e["tstep1"])) + 1)]
    steps += [(i, e["tstep2"], e["tdiv"] + i * e["tstep2"]) for i in range(1, int(round((e["ttot"] - e["tdiv"]) / e["tstep2"])) + 1)]
    ft = f0
    for step, (_, ht, t2) in enumerate(steps, start=1):
        if max_steps is not None and step > max_steps:
            break
        # exponential transformation (%EXTR, mod_trans.f90:127-135,150-154): not in the very first step
        if p.bextr and step > 1:
            omeg = np.asfortranarray(np.log(f0 / ft) / ht)
        else:
            omeg = np.zeros((p.nnod, p.ng), order="F")
        # rod bank changes (mod_trans.f90:374-388)
        for b in range(c["nb"]):
            if mdir[b] == 1 and t2 - tmove[b] > 1e-5 and fbpos[b] - bpos[b] < 1e-5:
                bpos[b] = max(bpos[b] - ht * bspeed[b], fbpos[b])
            elif mdir[b] == 2 and t2 - tmove[b] > 1e-5 and fbpos[b] - bpos[b] > 1e-5:
                bpos[b] = min(bpos[b] + ht * bspeed[b], fbpos[b])
        p.update_xs(bpos)
        sigrp = p.sigr.copy(order="F")
        sigr = p.sigr.copy(order="F")
        for g in range(p.ng):
            sigr[:, g] = sigr[:, g] + 1.0 / (p.sth * velo[g] * ht) + omeg[:, g] / velo[g]
        ft, fst = f0.copy(order="F"), fs0.copy()
        _push_xs(solver, p, sigr=sigr)
        solver.set_transient(c0=c0, ft=ft, fst=fst, omeg=omeg, sigrp=sigrp, L=L)
        rc, maxi, n = solver.outer_tr(ht)
        assert rc == 0, rc
        st = solver.state()
        f0, fs0 = st["f0"], st["fs0"]
        # uPden (mod_trans.f90:601-644)
        for i in range(6):
            pxe = np.exp(-lamb[i] * ht)
            a1 = (1.0 - pxe) / (lamb[i] * ht)
            a2 = 1.0 - a1
            a1 = a1 - pxe
            c0[:, i] = c0[:, i] * pxe + ibeta[i] / lamb[i] * (a1 * fst + a2 * fs0)
        tpow2 = powtot(p, f0)
        L = _leakage(p, solver, f0)
        rho = reactivity(p, af, sigrp, f0, fs0, L)
        trace.append((step, t2, rho / ctbeta, tpow2 / tpow1, n, maxi))
        say(f"{step:4d} {t2:10.3f} {rho / ctbeta:10.4f} {tpow2 / tpow1:15.4E}  outers {n}")
    return trace


# ------------------------------------------------------------------------------------------------
# rod ejection with thermal-hydraulic feedback (mod_trans.f90:163-330 rod_eject_th, trans_calc with
# thc = 1): the NEACRP decks smpl/transient/NEACRP/A1t ... C2t
# ------------------------------------------------------------------------------------------------
def _time_steps(e):
    steps = [(i, e["tstep1"], i * e["tstep1"]) for i in range(1, int(round(e["tdiv"] / e["tstep1"])) + 1)]
    steps += [(i, e["tstep2"], e["tdiv"] + i * e["tstep2"]) for i in range(1, int(round((e["ttot"] - e["tdiv"]) / e["tstep2"])) + 1)]
    return steps


def _move_rods(e, bpos, mdir, ht, t2):
    """rod bank changes (mod_trans.f90:374-388)"""
    fbpos, tmove, bspeed = e["fbpos"], e["tmove"], e["bspeed"]
    for b in range(len(bpos)):
        if mdir[b] == 1 and t2 - tmove[b] > 1e-5 and fbpos[b] - bpos[b] < 1e-5:
            bpos[b] = max(bpos[b] - ht * bspeed[b], fbpos[b])
        elif mdir[b] == 2 and t2 - tmove[b] > 1e-5 and fbpos[b] - bpos[b] > 1e-5:
            bpos[b] = min(bpos[b] + ht * bspeed[b], fbpos[b])


def _kinetics(p):
    """Kinetics data of a RODEJECT deck.  %XSEC decks: one set on the %EJCT card.  %XTAB decks
    (bxtab = 1): one set per material in the library, m(mat)%iBeta / %lamb / %velo.  Returns
    (xtab?, ibeta, lamb, velo, tbeta): (6,) / (ng,) arrays, or (nmat, 6) / (nmat, ng) for %XTAB;
    tbeta(nmat) accumulated family by family (mod_trans.f90:235-249)."""
    if getattr(p, "xtab", None) is None:
        e = p.ejct
        tbeta = np.full(p.nmat, 0.0)
        for jf in range(6):
            tbeta = tbeta + e["ibeta"][jf]
        return False, e["ibeta"], e["lamb"], e["velo"], tbeta
    ibeta = np.array([t["ibeta"] for t in p.xtab])
    lamb = np.array([t["lamb"] for t in p.xtab])
    velo = np.array([t["velo"] for t in p.xtab])
    tbeta = np.zeros(p.nmat)
    for jf in range(6):
        tbeta = tbeta + ibeta[:, jf]
    return True, ibeta, lamb, velo, tbeta


def calc_beta(p, af, f0, nuf, ibeta):
    """calc_beta (mod_trans.f90:692-741): adjoint-weighted core-averaged delayed neutron fraction
    (%XTAB decks only)."""
    m = p.mat - 1
    vdum = np.zeros(p.nnod)
    for g in range(p.ng):
        vdum = vdum + nuf[:, g] * f0[:, g]
    vdum2 = np.zeros(p.nnod)
    for g in range(p.ng):
        vdum2 = vdum2 + p.chi[m, g] * vdum * af[:, g]
    F = _integrate(p, vdum2)
    ctbeta = 0.0
    for i in range(6):
        vdum2 = np.zeros(p.nnod)
        for g in range(p.ng):
            vdum2 = vdum2 + p.chi[m, g] * ibeta[m, i] * vdum * af[:, g]
        ctbeta = ctbeta + _integrate(p, vdum2) / F
    return ctbeta


def _integrate(p, s):
    """Integrate (mod_cmfd.f90:1120-1139): serial sum of vdel * s"""
    tot = 0.0
    for v in (p.vdel * s).tolist():
        tot = tot + v
    return tot


def rod_eject_th(p, g, max_steps=None, log=None):
    """rod_eject_th + trans_calc(thc = 1) on a thermal.HostGlue `g` (numpy glue; g.s = oracle.Oracle or
    capi.Solver, g.thm = the th module).  Returns rows (step, t, reactivity [$], relative power
    xppow, outer iterations, maxi, max fuel-centreline temperature [K])."""
    from . import thermal
    s, thm, th = g.s, g.thm, g.th
    e, c = p.ejct, p.crod
    xt, ibeta, lamb, velo, tbeta = _kinetics(p)
    bcon = p.fbk["bcon"]["val"]
    g.bpos = c["bpos"].astype(np.float64).copy()
    mdir = np.where(np.abs(e["fbpos"] - g.bpos) < 1e-5, 0, np.where(e["fbpos"] - g.bpos > 1e-5, 2, 1))
    thermal.th_iter(g, bcon, ind=0)
    ke = s.state()["Ke"]
    if abs(ke - 1.0) > 1e-5 and not xt:                       # KNE1 (mod_trans.f90:483-518; not for %XTAB decks, :209)
        for it in range(10):
            p.xnuf = p.xnuf / ke
            c["dnuf"] = c["dnuf"] / ke
            g.xs_update(bcon)
            s.outer(0)
            ke = s.state()["Ke"]
            if abs(ke - 1.0) < 1e-5:
                break
    s.outer_ad(0)
    af = s.state()["f0"].copy()
    s.outer(0)
    st = s.state()
    f0, fs0 = st["f0"], st["fs0"]
    tpow1 = powtot(p, f0)
    m = p.mat - 1
    if xt:                                                    # iPden, bxtab = 1: precursors only where nuf(n, ng) > 0
        fuel = p.nuf[:, p.ng - 1] > 0.0
        with np.errstate(divide="ignore", invalid="ignore"):  # the reflector's lamb = 0 never enters (fuel only)
            c0 = np.asfortranarray(np.where(fuel[:, None], (ibeta / lamb)[m, :] * fs0[:, None], 0.0))
        ctbeta = calc_beta(p, af, f0, p.nuf, ibeta)
    else:
        c0 = np.asfortranarray((ibeta / lamb)[None, :] * fs0[:, None])
        ctbeta = tbeta[0]
    L = _leakage(p, s, f0)
    rho = reactivity(p, af, p.sigr, f0, fs0, L)
    trace = [(0, 0.0, rho / ctbeta, th["ppow"] * 0.01, 0, False, float(g.st["tfm"][:, 0].max()))]
    if xt:
        s.set_kinetics_xtab(ibeta, lamb, velo, tbeta, p.sth, p.bth)
    else:
        s.set_kinetics(ibeta, lamb, velo, tbeta, p.sth, p.bth)
    ft = f0
    for step, (_, ht, t2) in enumerate(_time_steps(e), start=1):
        if max_steps is not None and step > max_steps:
            break
        omeg = np.asfortranarray(np.log(f0 / ft) / ht) if (p.bextr and step > 1) else np.zeros((p.nnod, p.ng), order="F")
        _move_rods(e, g.bpos, mdir, ht, t2)
        p.update_xs(g.bpos, bcon=bcon, ftem=g.ftem, mtem=g.mtem, cden=g.cden)
        sigrp = p.sigr.copy(order="F")
        sigr = p.sigr.copy(order="F")
        for gg in range(p.ng):
            vg = velo[m, gg] if xt else velo[gg]               # m(mat(n))%velo(g) for %XTAB decks (mod_trans.f90:405-411)
            sigr[:, gg] = sigr[:, gg] + 1.0 / (p.sth * vg * ht) + omeg[:, gg] / vg
        ft, fst = f0.copy(order="F"), fs0.copy()
        _push_xs(s, p, sigr=sigr)
        s.set_transient(c0=c0, ft=ft, fst=fst, omeg=omeg, sigrp=sigrp, L=L)
        rc, maxi, n = s.outer_tr(ht)
        assert rc == 0, rc
        st = s.state()
        f0, fs0 = st["f0"], st["fs0"]
        for i in range(6):                                    # uPden
            if xt:
                fuel = p.nuf[:, p.ng - 1] > 0.0
                lam, bet = lamb[m, i], ibeta[m, i]
            else:
                fuel, lam, bet = slice(None), lamb[i], ibeta[i]
            with np.errstate(divide="ignore", invalid="ignore"):
                pxe = np.exp(-lam * ht)
                a1 = (1.0 - pxe) / (lam * ht)
                a2 = 1.0 - a1
                a1 = a1 - pxe
                new = c0[:, i] * pxe + bet / lam * (a1 * fst + a2 * fs0)
            c0[fuel, i] = new[fuel]
        tpow2 = powtot(p, f0)
        L = _leakage(p, s, f0)
        rho = reactivity(p, af, sigrp, f0, fs0, L)
        rc, npow = s.powdis()
        xppow = th["ppow"] * tpow2 / tpow1 * 0.01
        nf = th["node_nf"][p.ix - 1, p.iy - 1]
        pline = npow * th["pow"] * xppow / (nf * p.zdel[p.iz - 1])
        thm.th_trans(p, th, g.st, pline, ht)
        g.ftem, g.mtem, g.cden = g.st["ftem"], g.st["mtem"], g.st["cden"]
        trace.append((step, t2, rho / ctbeta, xppow, n, maxi, float(g.st["tfm"][:, 0].max())))
        if log:
            log(f"{step:4d} {t2:9.4f} {rho / ctbeta:10.4f} {xppow:13.5E}  outers {n}  Tf,max {trace[-1][6]:.2f}")
    return trace


def rod_eject_th_device(p, g, max_steps=None, log=None):
    """The same transient on a thermal.DeviceGlue `g`: XS update with feedback, time-step glue,
    outer_tr, uPden, PowTot, reactivity, PowDis -> pline and th_trans all on the device; per step the
    bank positions go up and (reactivity, power, max fuel temperature on request) come back."""
    from . import thermal
    s, th = g.s, g.th
    e, c = p.ejct, p.crod
    xt, ibeta, lamb, velo, tbeta = _kinetics(p)
    bcon = p.fbk["bcon"]["val"]
    g.bpos = c["bpos"].astype(np.float64).copy()
    mdir = np.where(np.abs(e["fbpos"] - g.bpos) < 1e-5, 0, np.where(e["fbpos"] - g.bpos > 1e-5, 2, 1))
    thermal.th_iter(g, bcon, ind=0)
    ke = s.state()["Ke"]
    if abs(ke - 1.0) > 1e-5 and not xt:
        for it in range(10):
            p.xnuf = p.xnuf / ke
            c["dnuf"] = c["dnuf"] / ke
            s.set_material_xs(p)
            s.set_crod(p)
            g.xs_update(bcon)
            s.outer(0)
            ke = s.state()["Ke"]
            if abs(ke - 1.0) < 1e-5:
                break
    s.outer_ad(0)
    s.save_adjoint()
    af = s.state()["f0"] if xt else None
    s.outer(0)
    if xt:
        # calc_beta: a printed, once-per-run quantity (reactivity in $) -- evaluated on the host from the
        # adjoint and forward flux that come back once
        ctbeta = calc_beta(p, af, s.state()["f0"], s.get_xs()["nuf"], ibeta)
        s.set_kinetics_xtab(ibeta, lamb, velo, tbeta, p.sth, p.bth)
    else:
        ctbeta = tbeta[0]
        s.set_kinetics(ibeta, lamb, velo, tbeta, p.sth, p.bth)
    tpow1 = s.powtot()
    s.ipden()
    rho = s.reactivity(0)
    trace = [(0, 0.0, rho / ctbeta, th["ppow"] * 0.01, 0, False, float(s.th_state()["tfm"][:, 0].max()))]
    for step, (_, ht, t2) in enumerate(_time_steps(e), start=1):
        if max_steps is not None and step > max_steps:
            break
        _move_rods(e, g.bpos, mdir, ht, t2)
        s.update_omeg(ht, p.bextr and step > 1)
        g.xs_update(bcon)
        s.begin_time_step(ht)
        rc, maxi, n = s.outer_tr(ht)
        assert rc == 0, rc
        s.upden(ht)
        tpow2 = s.powtot()
        rho = s.reactivity(1)
        xppow = th["ppow"] * tpow2 / tpow1 * 0.01
        s.th_pline(th["pow"], xppow, form=1)
        rc = s.th_trans(None, ht)
        if rc > 0:
            raise thermal.StopError(s.last_error())
        tmax = float(s.th_state()["tfm"][:, 0].max())       # the printed Max. Tf (par_max(tfm(:,1)), mod_trans.f90:452)
        trace.append((step, t2, rho / ctbeta, xppow, n, maxi, tmax))
        if log:
            log(f"{step:4d} {t2:9.4f} {rho / ctbeta:10.4f} {xppow:13.5E}  outers {n}  Tf,max {tmax:.2f}")
    return trace
