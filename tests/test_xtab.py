"""%XTAB branch-table cross sections (inp_xtab / XStab_updt / brInterp / crod_tab_updt,
mod_io.f90:3648-4061, mod_xsec.f90:50-86,300-390,520-788) on the MOX/UO2 benchmark decks of the reference
(smpl/static/MOX/part2_*, part3_*).

Pinned by the reference's own numbers: smpl/transient/MOX/part4_<library> starts from the critical boron
concentration ADPRES found for part 3 with the same library (%BCON: 1341.99 ppm HELIOS, 1207.06 ppm
SERPENT); the oracle reproduces both through this path (library parsing, interpolation, rodded tables,
ADFs from the tables, TH feedback, boron search).

CPU tests: the numpy restatement (adpres_b200/deck.py) against a scalar transcription of brInterp and
against the golden values; the per-node DEVICE code (csrc/xtab_node.cuh, what k_xs_update_xtab runs per
thread) compiled for the host with g++ and compared bit for bit with the numpy restatement.  GPU tests:
adp_xs_update_xtab bit-exact against numpy, and the critical boron search device-resident."""
import ctypes as C
import json
import os
import shutil
import subprocess

import numpy as np
import pytest

from conftest import GOLDEN, ROOT, load_problem


def _gold(key):
    return json.load(open(os.path.join(GOLDEN, "mox_bcon.json")))["ppm"][key]


def _random_fields(p, rng, wide=False):
    """TH fields inside the tables of every library (a little outside with wide=True: the 20 % rule)"""
    n = p.nnod
    lo, hi = (0.56, 0.85) if wide else (0.662, 0.752)
    return (500.0 + 900.0 * rng.random(n) if wide else 560.0 + 760.0 * rng.random(n),      # ftem (tables: 560 .. 1320 K)
            540.0 + 60.0 * rng.random(n),                                                  # mtem (single branch: unused)
            lo + (hi - lo) * rng.random(n))                                                # cden (tables: 0.661 .. 0.752)


def test_xtab_library_parsing():
    p = load_problem("MOX_P3_HELIOS")
    assert (p.mode, p.ng, p.nmat, p.nnod) == ("BCSEARCH", 2, 18, 5654)
    assert p.crod is not None and p.crod["nb"] == 8 and p.ther is not None and p.fbk is None
    t = p.xtab[0]                                          # u42, 0.15 GWd/t
    assert (t["tadf"], t["trod"], t["nd"], t["nb"], t["nf"], t["nm"]) == (1, 1, 3, 3, 3, 1)
    assert np.array_equal(t["pd"], [0.66114, 0.71187, 0.75206]) and np.array_equal(t["pb"], [0.0, 1000.0, 2000.0])
    assert np.array_equal(t["pf"], [560.0, 900.0, 1320.0]) and t["pm"].shape == (1,)
    # first and last transport numbers of the first table of smpl/xsec/HELIOS/2G_XSEC_u42 (group 1; lines run over
    # density, then boron, fuel temperature)
    assert t["xs"][0, 0, 0, 0, 0] == 2.29048E-01 and t["xs"][2, 0, 0, 0, 0] == 2.44170E-01 and t["xs"][2, 2, 2, 0, 0] == 2.40836E-01
    assert t["xs"].shape == (3, 3, 3, 1, 24) and t["rxs"].shape == (3, 3, 3, 1, 24)
    assert t["rxs"][0, 0, 0, 0, 0] == 2.22040E-01
    # one ADF per group copied to the six faces (tadf = 1)
    assert np.all(t["xs"][..., 12:18] == t["xs"][..., 12:13]) and np.all(t["xs"][..., 18:24] == t["xs"][..., 18:19])
    assert np.array_equal(p.chi[0], [1.0, 0.0]) and abs(1.0 / t["velo"][1] - 2.36914E-06) < 1e-9
    # the compositions of one file differ (burn-up points), the reflector has a single density / temperature branch
    assert not np.array_equal(p.xtab[0]["xs"], p.xtab[1]["xs"])
    r = p.xtab[17]
    assert (r["nd"], r["nb"], r["nf"], r["nm"]) == (1, 3, 1, 1) and r["pd"][0] == 0.0
    # defaults of inp_ther for %XTAB decks (single-precision literals)
    d = p.xtab_defaults()
    assert d == dict(bcon=0.0, ftem=900.0, mtem=500.0, cden=float(np.float32(0.711)))


def _br_interp_scalar(t, rod, xcden, xbcon, xftem, xmtem):
    """brInterp transcribed statement by statement for ONE node (mod_xsec.f90:520-788)"""
    tab = t["rxs"] if rod else t["xs"]

    def closest(x, par, dim, absolute):
        i1 = i2 = 1
        if dim > 1:
            mx = dim
            if x >= par[0] and x <= par[mx - 1]:
                for s in range(2, mx + 1):
                    if x >= par[s - 2] and x <= par[s - 1]:
                        i1, i2 = s - 1, s
                        break
            elif x < par[0] and ((par[0] - x) < 100.0 if absolute else (par[0] - x) / par[0] < float(np.float32(0.2))):
                i1, i2 = 1, 2
            elif x > par[mx - 1] and ((x - par[mx - 1]) < 100.0 if absolute else (x - par[mx - 1]) / par[mx - 1] < float(np.float32(0.2))):
                i1, i2 = mx - 1, mx
            else:
                raise ValueError("out of range")
        return i1 - 1, i2 - 1
    s1, s2 = closest(xcden, t["pd"], t["nd"], False)
    t1, t2 = closest(xbcon, t["pb"], t["nb"], True)
    u1, u2 = closest(xftem, t["pf"], t["nf"], False)
    v1, v2 = closest(xmtem, t["pm"], t["nm"], False)
    X = tab
    if t["nm"] > 1:
        radx = (xmtem - t["pm"][v1]) / (t["pm"][v2] - t["pm"][v1])
        f = lambda a, b, c: X[a, b, c, v1] + radx * (X[a, b, c, v2] - X[a, b, c, v1])
    else:
        f = lambda a, b, c: X[a, b, c, v1].copy()
    xs = [None, f(s1, t1, u1), f(s1, t1, u2), f(s1, t2, u1), f(s1, t2, u2), f(s2, t1, u1), f(s2, t1, u2), f(s2, t2, u1), f(s2, t2, u2)]
    if t["nf"] > 1:
        radx = (xftem - t["pf"][u1]) / (t["pf"][u2] - t["pf"][u1])
        for i in (1, 3, 5, 7):
            xs[i] = xs[i] + radx * (xs[i + 1] - xs[i])
    if t["nb"] > 1:
        radx = (xbcon - t["pb"][t1]) / (t["pb"][t2] - t["pb"][t1])
        xs[1] = xs[1] + radx * (xs[3] - xs[1])
        xs[5] = xs[5] + radx * (xs[7] - xs[5])
    if t["nd"] > 1:
        xs[1] = xs[1] + (xcden - t["pd"][s1]) / (t["pd"][s2] - t["pd"][s1]) * (xs[5] - xs[1])
    return xs[1]


def test_vectorised_interpolation_equals_scalar_transcription():
    p = load_problem("MOX_P3_SERPENT")
    rng = np.random.default_rng(5)
    ftem, mtem, cden = _random_fields(p, rng, wide=True)
    for bcon in (0.0, 1207.06, 2050.0, -60.0):
        for mn in (0, 6, 11, 14, 17):
            t = p.xtab[mn]
            sel = rng.choice(p.nnod, 40, replace=False)
            for rod in ((0, 1) if t["trod"] == 1 else (0,)):
                v = p._br_interp(t, rod, cden[sel], bcon, ftem[sel], mtem[sel])
                for k, n in enumerate(sel):
                    assert np.array_equal(v[k], _br_interp_scalar(t, rod, cden[n], bcon, ftem[n], mtem[n]))
    # on a branch point the table value itself comes out; half way between two the mean
    t = p.xtab[0]
    v = p._br_interp(t, 0, np.array([t["pd"][1]]), 1000.0, np.array([900.0]), np.array([500.0]))
    assert np.array_equal(v[0], t["xs"][1, 1, 1, 0])
    v = p._br_interp(t, 0, np.array([t["pd"][0]]), 500.0, np.array([560.0]), np.array([500.0]))
    assert np.allclose(v[0], 0.5 * (t["xs"][0, 0, 0, 0] + t["xs"][0, 1, 0, 0]), rtol=1e-15)
    # beyond 20 % (boron: 100 ppm) outside the tables the reference STOPs
    for kw in (dict(cden=np.array([0.5])), dict(bcon=2100.0), dict(ftem=np.array([400.0])), dict(bcon=-100.0)):
        a = dict(cden=np.array([0.7]), bcon=1000.0, ftem=np.array([900.0]), mtem=np.array([500.0]))
        a.update(kw)
        with pytest.raises(ValueError, match="OUT OF THE RANGE"):
            p._br_interp(t, 0, a["cden"], a["bcon"], a["ftem"], a["mtem"])


def test_rodded_nodes_and_suppression():
    p = load_problem("MOX_P3_HELIOS")
    ftem, mtem, cden = np.full(p.nnod, 560.0), np.full(p.nnod, 500.0), np.full(p.nnod, 0.7518)
    # banks 1-4 at step 0 (fully inserted), 5-8 at 200: the tip of the withdrawn banks is 21.42 + 365.76 cm above the
    # core bottom = the upper edge of the fuel, so they only cover the top reflector node
    p.update_xs(p.crod["bpos"], bcon=1341.99, ftem=ftem, mtem=mtem, cden=cden)
    rodded = {k: getattr(p, k).copy() for k in ("sigtr", "siga", "nuf", "sigf", "sigs", "dc", "D", "sigr")}
    w = p.rod_fractions(p.crod["bpos"])
    cols = p.rodded_columns()
    top = p.iz == p.nzz
    assert np.all(w[cols & top] >= 0.0) and np.all(w[~cols] == -1.0)
    p.update_xs(np.full(8, 200.0), bcon=1341.99, ftem=ftem, mtem=mtem, cden=cden)     # all banks withdrawn to step 200
    inner = (w == 1.0) & (p.iz > 1) & (p.iz < p.nzz)
    assert inner.sum() > 0
    assert np.all(rodded["siga"][inner, 1] > p.siga[inner, 1])            # rods absorb thermal neutrons
    same = w < 0.0
    for k in ("sigtr", "siga", "nuf", "sigf", "sigs", "dc"):
        assert np.array_equal(rodded[k][same], getattr(p, k)[same]), k
    # a partially inserted bank: the node with the tip is the volume-weighted mix
    bpos = np.array([100.5, 0.0, 0.0, 0.0, 200.0, 200.0, 200.0, 200.0])
    w = p.rod_fractions(bpos)
    part = (w > 0.0) & (w < 1.0)
    assert part.sum() > 0
    p.update_xs(bpos, bcon=1341.99, ftem=ftem, mtem=mtem, cden=cden)
    n = np.nonzero(part)[0][0]
    t = p.xtab[p.mat[n] - 1]
    un = _br_interp_scalar(t, 0, cden[n], 1341.99, ftem[n], mtem[n])
    ro = _br_interp_scalar(t, 1, cden[n], 1341.99, ftem[n], mtem[n])
    mix = (1.0 - w[n]) * un + w[n] * ro
    assert np.array_equal(p.sigtr[n], mix[0:2]) and np.array_equal(p.siga[n], mix[2:4])
    assert np.array_equal(p.dc[n], mix[12:24].reshape(2, 6))
    assert np.array_equal(p.D[n], 1.0 / (3.0 * mix[0:2])) and p.sigr[n, 0] == mix[2] + mix[9] and p.sigr[n, 1] == mix[3] + mix[10]


@pytest.mark.parametrize("case", ["P3_HELIOS", "P3_SERPENT"])
def test_oracle_reproduces_the_reference_critical_boron_of_mox_part3(case):
    """The golden pin of this path: hot zero power, banks 1-4 inserted, critical boron search with the TH
    loop running -- the reference's own result is on the %BCON card of part 4."""
    from adpres_b200 import thermal
    from oracle import Oracle, th as oth
    p = load_problem("MOX_" + case)
    g = thermal.HostGlue(p, Oracle(p), oth)
    bc, rows = thermal.cbsearcht(g)
    assert rows[0][1] == 0.0 and rows[1][1] == 500.0               # rbcon is never set for %XTAB decks -> 0, then 500
    assert abs(bc - _gold(case)) < 0.02, (case, bc)


def test_oracle_mox_part2_hot_full_power():
    """Part 2 (hot full power, all rods out, no reference-produced number): PARCS 2-group nodal in the
    benchmark report finds about 1 680 ppm; loose sanity bounds only."""
    from adpres_b200 import thermal
    from oracle import Oracle, th as oth
    p = load_problem("MOX_P2_HELIOS")
    g = thermal.HostGlue(p, Oracle(p), oth)
    bc, rows = thermal.cbsearcht(g)
    assert 1650.0 < bc < 1720.0 and abs(rows[-1][2] - 1.0) < 1e-5
    f = g.th_fields()
    fuel = p.nuf[:, 1] > 0
    assert 800.0 < f["ftem"][fuel].mean() < 900.0 and 590.0 < f["mtem"].max() < 617.0


# ------------------------------------------------------------------ the device code, compiled for the host
@pytest.fixture(scope="module")
def hostlib(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("hostcheck") / "libxtab_host.so")
    src = os.path.join(ROOT, "tests", "hostcheck", "xtab_host.cpp")
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Werror", src, "-o", out],
                   check=True, capture_output=True)
    return C.CDLL(out)


def _host_update(lib, p, bcon, ftem, mtem, cden, bpos):
    from adpres_b200.capi import pack_xtab, _d, _ip
    dims, trod, par, xs, rxs = pack_xtab(p)
    N, G = p.nnod, p.ng
    out = dict(D=np.zeros((N, G), order="F"), sigr=np.zeros((N, G), order="F"), nuf=np.zeros((N, G), order="F"),
               sigf=np.zeros((N, G), order="F"), sigs=np.zeros((N, G, G), order="F"), dc=np.zeros((N, G, 6), order="F"))
    w = np.zeros(N)
    fb = bp = None
    if p.crod is not None:
        ia, ja, _ = p._node_assembly_maps()
        fbmap = p.crod["bmap"][np.ix_(ia, ja)]
        fb = np.ascontiguousarray(fbmap[p.ix[:p.npl] - 1, p.iy[:p.npl] - 1].astype(np.int32))
        bp = np.ascontiguousarray(bpos, dtype=np.float64)
    mat = np.ascontiguousarray(p.mat.astype(np.int32))
    rc = lib.xtab_host_update(G, p.nmat, dims.ctypes.data_as(_ip), trod.ctypes.data_as(_ip), _d(par), _d(xs), _d(rxs),
                              p.npl, p.nzz, mat.ctypes.data_as(_ip), None if fb is None else fb.ctypes.data_as(_ip), _d(bp),
                              _d(np.ascontiguousarray(p.zdel)), C.c_double(0.0 if p.crod is None else p.crod["pos0"]),
                              C.c_double(0.0 if p.crod is None else p.crod["ssize"]), C.c_double(bcon),
                              _d(np.ascontiguousarray(ftem)), _d(np.ascontiguousarray(mtem)), _d(np.ascontiguousarray(cden)),
                              _d(out["D"]), _d(out["sigr"]), _d(out["nuf"]), _d(out["sigf"]), _d(out["sigs"]), _d(out["dc"]), _d(w))
    return rc, out, w


@pytest.mark.parametrize("name", ["MOX_P3_HELIOS", "MOX_P3_SERPENT", "MOX_P2_HELIOS"])
def test_device_node_code_on_the_host_is_bit_exact(hostlib, name):
    """csrc/xtab_node.cuh (the body of k_xs_update_xtab) compiled with g++ against the numpy XStab_updt:
    random TH fields (also in the extrapolation margins), several boron concentrations, rods inserted,
    withdrawn, with the tip inside a node, exactly on a node boundary and above the core."""
    p = load_problem(name)
    rng = np.random.default_rng(17)
    nb = 0 if p.crod is None else p.crod["nb"]
    z_edge = 21.42 + 3 * 18.288                     # a node boundary: tip exactly on it for step = (z_edge - pos0) / ssize
    cases = [(0.0, np.zeros(nb)), (1341.99, None), (2080.0, np.full(nb, 200.0)), (-40.0, np.full(nb, 100.5)),
             (777.7, np.full(nb, (z_edge - 21.42) / 1.8288)), (500.0, np.full(nb, 212.0))]
    for k, (bcon, bpos) in enumerate(cases):
        ftem, mtem, cden = _random_fields(p, rng, wide=(k % 2 == 1))
        if p.crod is not None and bpos is None:
            bpos = p.crod["bpos"]
        p.update_xs(bpos, bcon=bcon, ftem=ftem, mtem=mtem, cden=cden)
        rc, out, w = _host_update(hostlib, p, bcon, ftem, mtem, cden, bpos)
        assert rc == 0
        if p.crod is not None:
            assert np.array_equal(w, p.rod_fractions(bpos)), k
        for key in ("D", "sigr", "nuf", "sigf", "sigs", "dc"):
            assert np.array_equal(out[key], getattr(p, key)), (name, k, key, np.abs(out[key] - getattr(p, key)).max())


def test_device_node_code_stop_codes(hostlib):
    from adpres_b200 import capi
    p = load_problem("MOX_P3_HELIOS")
    n = p.nnod
    ftem, mtem, cden = np.full(n, 900.0), np.full(n, 500.0), np.full(n, 0.7)
    assert _host_update(hostlib, p, 1000.0, ftem, mtem, cden, p.crod["bpos"])[0] == 0
    assert _host_update(hostlib, p, 2100.0, ftem, mtem, cden, p.crod["bpos"])[0] == capi.STOP_XTAB_RANGE
    bad = cden.copy(); bad[1234] = 0.4
    assert _host_update(hostlib, p, 1000.0, ftem, mtem, bad, p.crod["bpos"])[0] == capi.STOP_XTAB_RANGE
    # check_xs / Dsigr_updt after the update: a table that extrapolates to a negative nu-fission cross section ...
    keep = p.xtab[0]["xs"].copy()
    p.xtab[0]["xs"][:, 2, :, :, 4:6] = -0.5 * np.abs(keep[:, 2, :, :, 4:6])          # nuf at 2000 ppm
    assert _host_update(hostlib, p, 1900.0, ftem, mtem, cden, p.crod["bpos"])[0] == capi.STOP_XS_CHECK
    with pytest.raises(ValueError, match="NU\\*FISSION XS IS NEGATIVE"):
        p.update_xs(p.crod["bpos"], bcon=1900.0, ftem=ftem, mtem=mtem, cden=cden)
    assert _host_update(hostlib, p, 500.0, ftem, mtem, cden, p.crod["bpos"])[0] == 0
    # ... or a vanishing transport cross section
    p.xtab[0]["xs"][...] = keep
    p.xtab[0]["xs"][..., 0] = 1.0e-6
    assert _host_update(hostlib, p, 500.0, ftem, mtem, cden, p.crod["bpos"])[0] == capi.STOP_XS_CHECK
    with pytest.raises(ValueError, match="Negative diffusion coefficient"):
        p.update_xs(p.crod["bpos"], bcon=500.0, ftem=ftem, mtem=mtem, cden=cden)
    p.xtab[0]["xs"][...] = keep
    for t in p.xtab:
        t["trod"] = 0
    assert _host_update(hostlib, p, 1000.0, ftem, mtem, cden, p.crod["bpos"])[0] == capi.STOP_XTAB_NOROD
    with pytest.raises(ValueError, match="DOES NOT HAVE CONTROL ROD DATA"):
        p.update_xs(p.crod["bpos"], bcon=1000.0, ftem=ftem, mtem=mtem, cden=cden)


@pytest.fixture(scope="module")
def kinlib(tmp_path_factory):
    gxx = shutil.which("g++")
    if not gxx:
        pytest.skip("no g++")
    out = str(tmp_path_factory.mktemp("hostcheck") / "libkin_host.so")
    src = os.path.join(ROOT, "tests", "hostcheck", "kinetics_host.cpp")
    subprocess.run([gxx, "-O2", "-std=c++17", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Werror", src, "-o", out],
                   check=True, capture_output=True)
    return C.CDLL(out)


def _kin_state(p, seed=3):
    """random but physical transient state on the part-3 core: (f0, fs0, c0, ft, fst, omeg, sigrp, L, s0 with one column)"""
    rng = np.random.default_rng(seed)
    N, G = p.nnod, p.ng
    fuel = p.nuf[:, G - 1] > 0
    f0 = np.asfortranarray(0.5 + rng.random((N, G)))
    fs0 = np.where(fuel, 0.2 + rng.random(N), 0.0)
    c0 = np.asfortranarray(np.where(fuel[:, None], rng.random((N, 6)), 0.0))
    ft = np.asfortranarray(f0 * (1.0 + 0.01 * rng.standard_normal((N, G))))
    fst = fs0 * (1.0 + 0.01 * rng.standard_normal(N))
    omeg = np.asfortranarray(5.0 * rng.standard_normal((N, G)))
    sigrp = np.asfortranarray(p.sigr * (1.0 + 0.01 * rng.random((N, G))))
    L = np.asfortranarray(0.01 * rng.standard_normal((N, G)))
    s0 = np.zeros((N, G), order="F")
    s0[:, G - 1] = 0.05 * rng.random(N)
    return f0, fs0, c0, ft, fst, omeg, sigrp, L, s0


def test_device_kinetics_node_code_against_the_c_oracle(kinlib):
    """csrc/kinetics_node.cuh (bxtab = 1 branches of iPden, uPden, get_exsrc and the time-absorption term: kinetics
    data per material from the %XTAB library, precursors only in fuel) compiled with g++ against the C oracle."""
    from adpres_b200 import transient
    from adpres_b200.capi import _d, _ip
    from oracle import Oracle
    p = load_problem("MOX_P3_HELIOS")
    N, G = p.nnod, p.ng
    xt, ibeta, lamb, velo, tbeta = transient._kinetics(p)
    assert xt and ibeta.shape == (p.nmat, 6) and velo.shape == (p.nmat, G) and np.all(tbeta[:17] > 0.002)
    sth, ht = 0.7, 0.002
    bth = (1.0 - sth) / sth
    f0, fs0, c0, ft, fst, omeg, sigrp, L, s0 = _kin_state(p)
    o = Oracle(p)
    o.set_kinetics_xtab(ibeta, lamb, velo, tbeta, sth, bth)
    mat = np.ascontiguousarray(p.mat.astype(np.int32))
    kin = (G, p.nmat, _d(np.ascontiguousarray(lamb)), _d(np.ascontiguousarray(ibeta)), _d(np.ascontiguousarray(velo)), C.c_longlong(N),
           mat.ctypes.data_as(_ip))
    nuf = np.asfortranarray(p.nuf)
    fuel = p.nuf[:, G - 1] > 0
    # iPden
    o.set_state(f0, fs0, 1.0)
    o.ipden()
    c_h = np.full((N, 6), np.nan, order="F")
    kinlib.kin_host_ipden(*kin, _d(nuf), _d(fs0), _d(c_h))
    assert np.array_equal(c_h, o.transient()["c0"]) and np.all(c_h[~fuel] == 0.0) and np.all(c_h[fuel] > 0.0)
    # uPden
    o.set_transient(c0=c0, ft=ft, fst=fst, omeg=omeg, sigrp=sigrp, L=L)
    o.upden(ht)
    c_h = c0.copy(order="F")
    kinlib.kin_host_upden(*kin, _d(nuf), C.c_double(ht), _d(fst), _d(fs0), _d(c_h))
    assert np.array_equal(c_h, o.transient()["c0"]) and np.array_equal(c_h[~fuel], c0[~fuel])
    # get_exsrc (s0: only the column of the last group swept is non-zero)
    o.set_transient(c0=c0)
    o.L.orc_set_s0(o.h, _d(s0))
    o.get_exsrc(ht)
    t = o.transient()
    ex_h, df_h = np.zeros((N, G), order="F"), np.zeros(N)
    kinlib.kin_host_exsrc(*kin, _d(nuf), C.c_double(ht), C.c_double(sth), C.c_double(bth), _d(c0), _d(fst), _d(tbeta),
                          _d(np.asfortranarray(p.chi)), _d(L), _d(sigrp), _d(ft), _d(np.ascontiguousarray(s0[:, G - 1])), G - 1,
                          _d(omeg), _d(ex_h), _d(df_h))
    assert np.array_equal(df_h, t["dfis"]) and np.array_equal(ex_h, t["exsrc"])
    # the reflector has no delayed neutrons: dfis = 0, and a fuel node sees its own material's data
    assert np.all(df_h[~fuel] == 0.0) and np.all(df_h[fuel] > 0.0) and len(np.unique(df_h[fuel])) > 5
    # time-absorption term of trans_calc with m(mat(n))%velo(g)
    sigr = np.asfortranarray(p.sigr.copy())
    sp = np.zeros((N, G), order="F")
    kinlib.kin_host_time_absorption(*kin, C.c_double(sth), C.c_double(ht), _d(omeg), _d(sigr), _d(sp))
    m = p.mat - 1
    assert np.array_equal(sp, p.sigr)
    for g in range(G):
        assert np.array_equal(sigr[:, g], p.sigr[:, g] + 1.0 / (sth * velo[m, g] * ht) + omeg[:, g] / velo[m, g])


def test_oracle_mox_part4_first_steps():
    """smpl/transient/MOX/part4_helios (rod ejection from hot zero power, %XTAB + %EXTR, 22 616 nodes): steady state,
    adjoint, core-averaged delayed neutron fraction and the first time steps with the CPU oracle."""
    from adpres_b200 import thermal, transient
    from oracle import Oracle, th as oth
    p = load_problem("MOX_P4_HELIOS")
    assert (p.mode, p.nnod, p.bextr, p.crod["nb"]) == ("RODEJECT", 22616, 1, 9) and p.ejct["ibeta"] is None
    g = thermal.HostGlue(p, Oracle(p), oth)
    tr = transient.rod_eject_th(p, g, max_steps=3)
    assert abs(g.s.state()["Ke"] - 1.0) < 1e-3                 # critical at the reference's own boron (1341.99 ppm); no KNE1 for %XTAB
    assert abs(tr[0][2]) < 0.02 and abs(tr[0][3] - 1.0e-6) < 1e-18           # reactivity ~ 0 $, 1e-4 % power
    assert 0.0 < tr[1][2] < tr[2][2] < tr[3][2] < 0.01           # bank 9 starts to move out: reactivity rises
    assert tr[3][3] > tr[1][3] > 1.0e-6


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("name", ["MOX_P3_HELIOS", "MOX_P3_SERPENT", "MOX_P2_HELIOS"])
def test_gpu_xtab_update_bit_exact(name):
    """adp_xs_update_xtab (k_xs_update_xtab) against the numpy XStab_updt: same cases as the host check above"""
    from adpres_b200 import capi
    p = load_problem(name)
    s = capi.Solver(p)
    s.set_xtab(p)
    if p.crod is not None:
        s.set_crod_map(p)
    rng = np.random.default_rng(17)
    nb = 0 if p.crod is None else p.crod["nb"]
    z_edge = 21.42 + 3 * 18.288
    cases = [(0.0, np.zeros(nb)), (1341.99, None), (2080.0, np.full(nb, 200.0)), (-40.0, np.full(nb, 100.5)),
             (777.7, np.full(nb, (z_edge - 21.42) / 1.8288)), (500.0, np.full(nb, 212.0))]
    for k, (bcon, bpos) in enumerate(cases):
        ftem, mtem, cden = _random_fields(p, rng, wide=(k % 2 == 1))
        if p.crod is not None and bpos is None:
            bpos = p.crod["bpos"]
        p.update_xs(bpos, bcon=bcon, ftem=ftem, mtem=mtem, cden=cden)
        assert s.xs_update_xtab(bcon, ftem, mtem, cden, bpos if p.crod is not None else None) == 0
        x = s.get_xs()
        x["dc"] = s.get_dc()
        for key in ("D", "sigr", "nuf", "sigf", "sigs", "dc"):
            assert np.array_equal(x[key], getattr(p, key)), (name, k, key, np.abs(x[key] - getattr(p, key)).max())
    # the reference's STOPs come back as codes
    n = p.nnod
    ftem, mtem, cden = np.full(n, 900.0), np.full(n, 500.0), np.full(n, 0.7)
    bp = p.crod["bpos"] if p.crod is not None else None
    assert s.xs_update_xtab(2100.0, ftem, mtem, cden, bp) == capi.STOP_XTAB_RANGE
    assert "OUT OF THE RANGE" in s.last_error()
    assert s.xs_update_xtab(1000.0, ftem, mtem, cden, bp) == 0


@pytest.mark.gpu
@pytest.mark.parametrize("case", ["P3_HELIOS", "P3_SERPENT"])
def test_gpu_reproduces_the_reference_critical_boron_of_mox_part3(case):
    """the golden values with the whole loop (XStab_updt, outer_th, PowDis, th_upd) on the device"""
    from adpres_b200 import capi, thermal
    p = load_problem("MOX_" + case)
    bc, rows = thermal.cbsearcht(thermal.DeviceGlue(p, capi.Solver(p)))
    assert abs(bc - _gold(case)) < 0.02, (case, bc)


@pytest.mark.gpu
def test_gpu_mox_part2_search_follows_the_oracle():
    """hot full power: strong TH feedback through the tables; the device-resident search lands on the oracle's boron"""
    from adpres_b200 import capi, thermal
    from oracle import Oracle, th as oth
    p1, p2 = load_problem("MOX_P2_HELIOS"), load_problem("MOX_P2_HELIOS")
    go = thermal.HostGlue(p1, Oracle(p1), oth)
    gd = thermal.DeviceGlue(p2, capi.Solver(p2))
    bo, ro = thermal.cbsearcht(go)
    bd, rd = thermal.cbsearcht(gd)
    assert abs(bo - bd) < 0.05, (bo, bd)
    for a, b in zip(ro[:3], rd[:3]):                        # same first guesses (0, 500, first secant step)
        assert abs(a[1] - b[1]) < 0.05 and abs(a[2] - b[2]) < 5e-6, (a, b)
    fo, fd = go.th_fields(), gd.th_fields()
    for k in ("ftem", "mtem", "cden"):
        assert np.abs(fo[k] - fd[k]).max() / np.abs(fo[k]).max() < 1e-5, k


@pytest.mark.gpu
def test_gpu_mox_part4_first_steps_device_resident():
    """MOX part 4 (rod ejection, %XTAB kinetics per material, %EXTR): device-resident time stepping against
    the oracle-driven run through the first steps of the rod withdrawal"""
    from adpres_b200 import capi, thermal, transient
    from oracle import Oracle, th as oth
    ps = []
    for _ in range(2):
        p = load_problem("MOX_P4_HELIOS")
        p.serc = p.ferc = 1e-9          # converged steps: the exit iteration must not depend on round-off
        p.nout = 5000
        ps.append(p)
    to = transient.rod_eject_th(ps[0], thermal.HostGlue(ps[0], Oracle(ps[0]), oth), max_steps=4)
    td = transient.rod_eject_th_device(ps[1], thermal.DeviceGlue(ps[1], capi.Solver(ps[1])), max_steps=4)
    assert len(to) == len(td) == 5
    for a, b in zip(to, td):
        assert abs(a[2] - b[2]) < 1e-4, (a, b)                       # reactivity [$]
        assert abs(a[3] / b[3] - 1.0) < 1e-4, (a, b)                 # relative power (north star: 1e-4)
        assert abs(a[6] / b[6] - 1.0) < 1e-6                         # max fuel centreline temperature


@pytest.mark.gpu
def test_gpu_xs_update_reports_the_check_xs_stop():
    """Dsigr_updt / check_xs follow every XS update in the reference (mod_xsec.f90:41,83,217): a rod increment that
    cancels the transport cross section must come back as ADP_STOP_XS_CHECK from the %XSEC device path as well."""
    from adpres_b200 import capi
    p = load_problem("LMW")
    s = capi.Solver(p)
    s.set_material_xs(); s.set_crod()
    bpos = np.array([100.0, 100.0])
    assert s.xs_update(bpos) == 0
    p.crod["dsigtr"] = -p.xsigtr.copy()
    s.set_crod(p)
    assert s.xs_update(bpos) == capi.STOP_XS_CHECK and "diffusion coefficient" in s.last_error()
    with pytest.raises(ValueError, match="Negative diffusion coefficient"):
        p.update_xs(bpos)
