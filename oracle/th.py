"""CPU oracle for the thermal-hydraulic channel solve th_upd / th_trans (reference: src/mod_th.f90).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the product (adp_th_upd, adp_th_trans in
adpres_b200/csrc/th.cu) never imports this.  Pinned (th_upd, with the XS feedback and the boron
search around it) by the six critical boron concentrations the reference's NEACRP transient decks
carry (tests/golden/neacrp_bcon.json, tests/test_th.py); th_trans is unpinned.  The functions restate the cited lines with the same operand order;
the node loop of the reference (n = 1..nnod, k-major) is vectorised over the nodes of one plane
and kept serial over the planes, which preserves the only dependence it has (the enthalpy /
flow rate a channel hands from plane k to plane k + 1 through entm(i,j), bfrate(i,j)).

`th` is the dict of adpres_b200.deck.Problem.th_setup(); `st` the state dict
    tfm (nnod, nt+1), heatf, ent, ftem, mtem, cden (nnod) [, frate (nnod)]
which the functions update in place.  A steam-table range violation (the reference STOPs) raises
SteamTableError.
"""
import numpy as np


class SteamTableError(RuntimeError):
    pass


def getent(th, t):
    """getent (mod_th.f90:219-258): enthalpy at temperature t."""
    stab, ntem = th["stab"], th["ntem"]
    if t < stab[0, 0] or t > stab[ntem - 1, 0]:
        raise SteamTableError("MODERATOR TEMP. IS OUT OF THE RANGE OF DATA IN THE STEAM TABLE")
    t2, ent2 = stab[0, 0], stab[0, 2]
    for i in range(1, ntem):
        t1, ent1 = t2, ent2
        t2, ent2 = stab[i, 0], stab[i, 2]
        if t1 <= t <= t2:
            return ent1 + (t - t1) / (t2 - t1) * (ent2 - ent1)
    raise AssertionError


def gettd(th, ent):
    """gettd (mod_th.f90:261-317) for a vector of enthalpies -> t, rho, Pr, kv, tc, R"""
    stab, ntem = th["stab"], th["ntem"]
    h = stab[:, 2]
    i1 = np.full(ent.shape, -1, dtype=np.int64)
    inside = (ent >= h[0]) & (ent <= h[ntem - 1])
    for i in range(ntem - 1, 0, -1):                      # the reference takes the FIRST matching i
        m = inside & (ent >= h[i - 1]) & (ent <= h[i])
        i1[m] = i - 1
    low = (ent < h[0]) & ((h[0] - ent) / h[0] < 0.1)
    high = (ent > h[ntem - 1]) & ((ent - h[ntem - 1]) / h[ntem - 1] < 0.1)
    i1[low] = 0
    i1[high] = ntem - 2
    if (i1 < 0).any():
        raise SteamTableError("ENTHALPY IS OUT OF THE RANGE IN THE STEAM TABLE")
    i2 = i1 + 1
    ratx = (ent - h[i1]) / (h[i2] - h[i1])
    col = lambda c: stab[i1, c] + ratx * (stab[i2, c] - stab[i1, c])
    R = 1000.0 * (stab[i2, 1] - stab[i1, 1]) / (h[i2] - h[i1])
    return col(0), col(1), col(3), col(4), col(5), R


def getkc(t):
    return 7.51 + 2.09e-2 * t - 1.45e-5 * (t * t) + 7.67e-9 * (t * t * t)       # t**2, t**3: repeated products


def getkf(t):
    return 1.05 + 2150.0 / (t - 73.15)


def getcpc(t):
    return 252.54 + 0.11474 * t


def getcpf(t):
    return 162.3 + 0.3038 * t - 2.391e-4 * (t * t) + 6.404e-8 * (t * t * t)


def geths(th, xden, Pr, kv, tc):
    """geths (mod_th.f90:413-436).  NB the reference CALLS it as geths(cden, Pr, kv, tcon) although
    the dummy arguments are (xden, tc, kv, Pr): inside, `tc` holds the Prandtl number and `Pr` the
    conductivity (mod_th.f90:636, 507).  Restated with the call's actual association."""
    tc_dummy, kv_dummy, Pr_dummy = Pr, kv, tc
    cvelo = th["cflow"] / (th["farea"] * xden * 1000.0)
    Re = cvelo * th["dh"] / (kv_dummy * 1.0e-6)
    Nu = 0.023 * (Pr_dummy ** 0.4) * (Re ** 0.8)
    return (tc_dummy / th["dh"]) * Nu


def tridia_solve(a, b, c, d):
    """TridiaSolve (mod_th.f90:380-409), columns = systems.  a, b, c, d: (nt+1, m)"""
    n = d.shape[0]
    c = c.copy(); d = d.copy()
    c[0] = c[0] / b[0]
    d[0] = d[0] / b[0]
    for i in range(1, n):
        c[i] = c[i] / (b[i] - a[i] * c[i - 1])
        d[i] = (d[i] - a[i] * d[i - 1]) / (b[i] - a[i] * c[i - 1])
    x = np.zeros_like(d)
    x[n - 1] = d[n - 1]
    for i in range(n - 2, -1, -1):
        x[i] = d[i] - c[i] * x[i + 1]
    return x


def _pin_system(th, st, sl, hs, pdens, h=None):
    """rows of the radial conduction system for the nodes `sl` (mod_th.f90:641-686 steady, 511-574 transient)"""
    nt, rpos, rdel, rf, rg, rc = th["nt"], th["rpos"], th["rdel"], th["rf"], th["rg"], th["rc"]
    tfm = st["tfm"][sl, :]                                  # (m, nt+1)
    m = tfm.shape[0]
    Hg = 1.0e4
    fdens, cdens = 10.412e3, 6.6e3
    a = np.zeros((nt + 1, m)); b = np.zeros((nt + 1, m)); c = np.zeros((nt + 1, m)); d = np.zeros((nt + 1, m))
    tr = h is not None
    # fuel centreline
    kt1 = getkf(tfm[:, 0]); kt2 = getkf(tfm[:, 1])
    kt = 2.0 * kt1 * kt2 / (kt1 + kt2)
    xc = kt * rpos[0] / rdel[0]
    if tr:
        eta = fdens * getcpf(tfm[:, 0]) * rpos[0] ** 2 / (2.0 * h)
        b[0] = xc + eta
        d[0] = pdens * 0.5 * rpos[0] ** 2 + eta * tfm[:, 0]
    else:
        b[0] = xc
        d[0] = pdens * 0.5 * rpos[0] ** 2
    c[0] = -xc
    for i in range(2, nt - 1):                              # Fortran i = 2 .. nt-2
        kt1 = kt2
        kt2 = getkf(tfm[:, i])
        kt = 2.0 * kt1 * kt2 / (kt1 + kt2)
        xa = xc
        xc = kt * rpos[i - 1] / rdel[i - 1]
        a[i - 1] = -xa
        c[i - 1] = -xc
        if tr:
            eta = fdens * getcpf(tfm[:, i - 1]) * (rpos[i - 1] ** 2 - rpos[i - 2] ** 2) / (2.0 * h)
            b[i - 1] = xa + xc + eta
            d[i - 1] = pdens * 0.5 * (rpos[i - 1] ** 2 - rpos[i - 2] ** 2) + eta * tfm[:, i - 1]
        else:
            b[i - 1] = xa + xc
            d[i - 1] = pdens * 0.5 * (rpos[i - 1] ** 2 - rpos[i - 2] ** 2)
    # fuel-gap interface (row nt-1)
    xa = xc
    xc = rg * Hg
    a[nt - 2] = -xa
    c[nt - 2] = -xc
    if tr:
        eta = fdens * getcpf(tfm[:, nt - 2]) * (rf ** 2 - rpos[nt - 3] ** 2) / (2.0 * h)
        b[nt - 2] = xa + xc + eta
        d[nt - 2] = pdens * 0.5 * (rf ** 2 - rpos[nt - 3] ** 2) + eta * tfm[:, nt - 2]
    else:
        b[nt - 2] = xa + xc
        d[nt - 2] = pdens * 0.5 * (rf ** 2 - rpos[nt - 3] ** 2)
    # gap-cladding interface (row nt)
    kt1 = getkc(tfm[:, nt - 1]); kt2 = getkc(tfm[:, nt])
    kt = 2.0 * kt1 * kt2 / (kt1 + kt2)
    xa = xc
    xc = kt * rpos[nt - 1] / rdel[nt - 1]
    a[nt - 1] = -xa
    c[nt - 1] = -xc
    if tr:
        eta = cdens * getcpc(tfm[:, nt - 1]) * (rpos[nt - 1] ** 2 - rg ** 2) / (2.0 * h)
        b[nt - 1] = xa + xc + eta
        d[nt - 1] = eta * tfm[:, nt - 1]
    else:
        b[nt - 1] = xa + xc
        d[nt - 1] = 0.0
    # cladding-coolant interface (row nt+1)
    xa = xc
    a[nt] = -xa
    if tr:
        eta = cdens * getcpc(tfm[:, nt]) * (rc ** 2 - rpos[nt - 1] ** 2) / (2.0 * h)
        xc = rc * hs
        b[nt] = xa + xc + eta
        d[nt] = rc * hs * st["mtem"][sl] + eta * tfm[:, nt]
    else:
        b[nt] = xa + hs * rc
        d[nt] = rc * hs * st["mtem"][sl]
    return a, b, c, d


def th_upd(p, th, st, xpline):
    """th_upd (mod_th.f90:594-699)."""
    npl, pi = p.npl, th["pi"]
    enti = getent(th, th["tin"])
    entm = np.zeros(npl)
    alp = 0.7
    for k in range(p.nzz):
        sl = slice(k * npl, (k + 1) * npl)
        cpline = st["heatf"][sl] * pi * th["dia"] + th["cf"] * xpline[sl] * 100.0
        zd = p.zdel[k] * 0.01
        below = enti if k == 0 else entm
        ent = below + 0.5 * cpline * zd / th["cflow"]
        st["ent"][sl] = ent
        t, rho, Pr, kv, tcon, _ = gettd(th, ent)
        st["mtem"][sl] = t
        st["cden"][sl] = rho
        entm = 2.0 * ent - below
        hs = geths(th, rho, Pr, kv, tcon)
        pdens = (1.0 - th["cf"]) * 100.0 * xpline[sl] / (pi * th["rf"] ** 2)
        a, b, c, d = _pin_system(th, st, sl, hs, pdens)
        x = tridia_solve(a, b, c, d)
        st["tfm"][sl, :] = x.T
        st["ftem"][sl] = (1.0 - alp) * x[0] + alp * x[th["nt"] - 2]
        st["heatf"][sl] = hs * (x[th["nt"]] - st["mtem"][sl])


def th_trans(p, th, st, xpline, h):
    """th_trans (mod_th.f90:440-591); st["frate"] = cflow before the first call (:482-486)."""
    npl, pi = p.npl, th["pi"]
    enti = getent(th, th["tin"])
    if st.get("frate") is None:
        st["frate"] = np.full(p.nnod, th["cflow"])
    entp = st["ent"].copy()
    entm = np.zeros(npl)
    bfrate = np.zeros(npl)
    alpha = 0.7
    for k in range(p.nzz):
        sl = slice(k * npl, (k + 1) * npl)
        mdens = st["cden"][sl] * 1000.0
        cpline = st["heatf"][sl] * pi * th["dia"] + th["cf"] * xpline[sl] * 100.0
        vol = th["farea"] * p.zdel[k] * 0.01
        fr = st["frate"][sl]
        eps = mdens * vol / h
        below = enti if k == 0 else entm
        ent = (cpline * p.zdel[k] * 0.01 + 2.0 * fr * below + eps * entp[sl]) / (eps + 2.0 * fr)
        st["ent"][sl] = ent
        t, rho, Pr, kv, tcon, R = gettd(th, ent)
        st["mtem"][sl] = t
        st["cden"][sl] = rho
        entm = 2.0 * ent - below
        fbelow = th["cflow"] if k == 0 else bfrate
        frn = fbelow - 0.5 * vol / h * R * (ent - entp[sl])
        st["frate"][sl] = frn
        bfrate = 2.0 * frn - fbelow
        hs = geths(th, rho, Pr, kv, tcon)
        pdens = 100.0 * xpline[sl] / (pi * th["rf"] ** 2)
        a, b, c, d = _pin_system(th, st, sl, hs, pdens, h=h)
        x = tridia_solve(a, b, c, d)
        st["tfm"][sl, :] = x.T
        st["ftem"][sl] = (1.0 - alpha) * x[0] + alpha * x[th["nt"] - 2]
        st["heatf"][sl] = hs * (x[th["nt"]] - st["mtem"][sl])


def abs_e(new, old):
    """AbsE (mod_th.f90:94-119)"""
    m = np.abs(new) > 1.0e-10
    return float(np.abs(new - old)[m].max()) if m.any() else 0.0


def pline_static(p, th, npow):
    """th_iter (mod_th.f90:61-64): linear power density [W/cm]"""
    nf = th["node_nf"][p.ix - 1, p.iy - 1]
    return npow * th["pow"] * th["ppow"] * 0.01 / (nf * p.zdel[p.iz - 1])


def initial_state(p, th):
    """inp_ther (mod_io.f90:3101-3113): tfm = 900, heatf = 0; ftem / mtem / cden as the caller sets them"""
    n = p.nnod
    return dict(tfm=np.full((n, th["nt"] + 1), 900.0, order="F"), heatf=np.zeros(n), ent=np.zeros(n),
                ftem=np.full(n, 900.0), mtem=np.full(n, 560.0), cden=np.full(n, 0.75), frate=None)
