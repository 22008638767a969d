"""The north star asks that "the same input decks (smpl/static, smpl/transient) run unchanged".  Every sample deck of
the reference parses (CPU, needs the reference tree), and the decks the other test files do not cover run here:
smpl/static/FDM-1D, CBCsearch (critical boron search WITHOUT thermal-hydraulic feedback, `cbsearch`), the pin-power
decks MOX/Part1b/aroA1 and Part1d/ariE7, and the SERPENT variants of MOX part 1.

The k-eff / boron numbers asserted for these decks are regression values of the oracle (no ADPRES-produced number
exists for them in the reference tree); the GPU tests compare the CUDA path with the oracle on the same decks.  For decks
run with the reference's loose iteration control only converged values are compared (the outer
count of an unconverged inner iteration depends on the summation order)."""
import glob
import os

import numpy as np
import pytest

from conftest import REFERENCE, load_problem

FORWARD = {"FDM_1D": 1.155939, "MOX_1B_A1": 1.061909, "MOX_1D_E7": 0.991738, "MOX_ARO_SERPENT": 1.054422, "MOX_ARI_SERPENT": 0.983328}


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present (GPU box)")
def test_every_sample_deck_of_the_reference_parses():
    from adpres_b200.deck import read_deck
    decks = [f for f in sorted(glob.glob(os.path.join(REFERENCE, "smpl", "static", "**"), recursive=True) +
                               glob.glob(os.path.join(REFERENCE, "smpl", "transient", "**"), recursive=True))
             if os.path.isfile(f) and "neacrp_" not in os.path.basename(f)]       # neacrp_*: card files included by FILE
    assert len(decks) == 54
    modes = {}
    for f in decks:
        p = read_deck(f)
        assert p.nnod > 0 and np.isfinite(p.D).all() and (p.D > 0).all() and (p.sigr > 0).all(), f
        modes[p.mode] = modes.get(p.mode, 0) + 1
    assert modes == {"FORWARD": 32, "BCSEARCH": 11, "ADJOINT": 1, "FIXEDSRC": 1, "RODEJECT": 9}


@pytest.mark.skipif(not os.path.isdir(REFERENCE), reason="reference tree not present (GPU box)")
def test_every_sample_deck_runs_end_to_end_with_the_oracle():
    """All 54 decks through examples/run_deck.py's mode dispatch (forward / adjoint / fixed source / boron search with
    and without TH / rod ejection with and without TH, %XSEC and %XTAB) with the CPU oracle as back end; transients:
    steady state, adjoint and the first two time steps.  Every eigenvalue solve converges, every search ends critical."""
    import importlib.util
    from adpres_b200 import thermal
    from adpres_b200.deck import read_deck
    from oracle import Oracle, th as oth
    from conftest import ROOT
    spec = importlib.util.spec_from_file_location("run_deck", os.path.join(ROOT, "examples", "run_deck.py"))
    run_deck = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(run_deck)
    decks = [f for f in sorted(glob.glob(os.path.join(REFERENCE, "smpl", "static", "**"), recursive=True) +
                               glob.glob(os.path.join(REFERENCE, "smpl", "transient", "**"), recursive=True))
             if os.path.isfile(f) and "neacrp_" not in os.path.basename(f)]
    for f in decks:
        p = read_deck(f)
        o = Oracle(p)
        glue = thermal.HostGlue(p, o, oth) if (p.mode == "BCSEARCH" or (p.mode == "RODEJECT" and p.ther is not None)) else None
        r = run_deck.run(p, o, glue, steps=2, log=lambda *a: None)
        if p.mode in ("FORWARD", "ADJOINT", "FIXEDSRC"):
            assert r["status"] == 0 and r["outers"] < p.nout, f
            assert p.mode == "FIXEDSRC" or 0.9 < r["keff"] < 1.2, (f, r["keff"])
        elif p.mode == "BCSEARCH":
            assert abs(r["keff"] - 1.0) < 1e-5 and 500.0 < r["bcon"] < 1800.0, (f, r["bcon"])
        else:
            assert len(r["trace"]) == 3 and all(np.isfinite(x[3]) and x[3] > 0 for x in r["trace"]), f


@pytest.mark.parametrize("deck", sorted(FORWARD))
def test_oracle_runs_the_remaining_forward_decks(deck):
    from oracle import Oracle
    p = load_problem(deck)
    o = Oracle(p)
    rc, n = o.outer(0)
    assert rc == 0 and n < p.nout
    assert abs(o.state()["Ke"] - FORWARD[deck]) < 2e-6, (deck, o.state()["Ke"])
    rc, pw = o.powdis()
    assert rc == 0 and abs(pw.sum() - 1.0) < 1e-12


def test_oracle_critical_boron_search_without_feedback():
    """smpl/static/CBCsearch: NEACRP A2 rod pattern, `cbsearch` (mod_th.f90:752-837): secant search on the boron
    concentration with a full `outer` solve per guess, cross sections at the reference conditions of the feedback cards."""
    from adpres_b200 import thermal
    from oracle import Oracle, th as oth
    p = load_problem("CBCsearch")
    assert p.mode == "BCSEARCH" and p.ther is None and p.fbk["bcon"]["ref"] == 1200.2
    g = thermal.HostGlue(p, Oracle(p), oth)
    bc, rows = thermal.cbsearch(g)
    assert rows[0][1] == 1200.2 and len(rows) <= 6 and abs(rows[-1][2] - 1.0) < 1e-5
    assert abs(bc - 1257.32) < 0.05, bc


# ------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("deck", sorted(FORWARD))
def test_gpu_remaining_forward_decks(deck):
    from adpres_b200 import capi
    from oracle import Oracle
    p = load_problem(deck)
    s, o = capi.Solver(p), Oracle(p)
    rc_s, n_s = s.outer(0)
    rc_o, n_o = o.outer(0)
    assert rc_s == rc_o == 0
    assert abs(s.state()["Ke"] - o.state()["Ke"]) * 1e5 < 1.0                    # north star: k-eff within 1 pcm
    _, pw_s = s.powdis()
    _, pw_o = o.powdis()
    nz = pw_o > 1e-12
    assert np.abs(pw_s[nz] / pw_o[nz] - 1.0).max() < 2e-4                        # converged to serc = ferc = 1e-5 each
    s.close()


@pytest.mark.gpu
def test_gpu_critical_boron_search_without_feedback():
    from adpres_b200 import capi, thermal
    from oracle import Oracle, th as oth
    p1, p2 = load_problem("CBCsearch"), load_problem("CBCsearch")
    bo, ro = thermal.cbsearch(thermal.HostGlue(p1, Oracle(p1), oth))
    bd, rd = thermal.cbsearch(thermal.DeviceGlue(p2, capi.Solver(p2)))
    assert abs(bo - bd) < 0.05, (bo, bd)
    assert abs(ro[0][2] - rd[0][2]) < 1e-5                                       # k-eff of the first guess (1200.2 ppm)


def _lmw_refined(rdiv, zdiv, tol):
    p = load_problem("LMW").refine(xdiv=[rdiv // 2] + [rdiv] * 5, ydiv=[rdiv // 2] + [rdiv] * 5, zdiv=[zdiv] * 10)
    p.nin, p.nupd, p.nac, p.nout, p.serc, p.ferc, p.biter = 10, 50, 5, 20000, tol, tol, 1
    return p


@pytest.mark.gpu
@pytest.mark.parametrize("fixture", ["lmw_refined_mid_oracle.json", "lmw_refined_full_oracle.json"])
def test_gpu_refined_lmw_transient_against_cpu_oracle_fixture(fixture):
    """BASELINE configs[4] (LMW rod ejection on a refined mesh) against the CPU oracle's trace, computed once by
    tools/lmw_refined_oracle.py and committed (mid: 2 cm x 2 cm x 4 cm, 146 250 nodes, 5 min of CPU; full: 1 cm x 1 cm x
    2 cm, 1.17 M nodes, the size tools/lmw_refined.py times; about an hour of CPU).  Converged far enough that the trace does
    not depend on the exit iteration (see the tolerance note below); device-resident time stepping (XS update, glue,
    outer_tr on the GPU)."""
    import json
    from conftest import GOLDEN
    from adpres_b200 import capi, transient
    path = os.path.join(GOLDEN, fixture)
    if not os.path.exists(path):
        pytest.skip("fixture not generated")
    fx = json.load(open(path))
    p = _lmw_refined(fx["rdiv"], fx["zdiv"], fx["serc"])
    assert p.nnod == fx["nnod"]
    s = capi.Solver(p)
    tr = transient.rod_eject_device_glue(p, s, max_steps=len(fx["trace"]) - 1, device_xs=True, step_tol=fx.get("step_tol"))
    assert len(tr) == len(fx["trace"])
    # mid fixture: every solve converged to 1e-8, the trace is independent of the exit iteration -> 1e-5.
    # full fixture (1 cm mesh): steady state / adjoint converged to 1e-8, time steps to 1e-7 (at 1e-8 the steps stall: each
    # nodal update perturbs the iterate at the 1e-6 level of the SANM constants).  Its own sensitivity, measured
    # with the oracle: steps at 1e-6 instead of 1e-7 (lmw_refined_full_oracle_t6.json) move the power by 2e-6 / 5e-6 / 2e-5 and
    # the reactivity by up to 1.6e-5 $ -- so 1e-7 steps are good for ~2e-6 and the north star's 1e-4 applies with margin.
    # (Round 1's fixture had the t = 0 state converged to 1e-6 only: its reactivity at t = 0 was 1.8e-5 $, the GPU's 3.8e-4 $,
    # the converged value is 1e-6 $ -- that difference was convergence noise of the steady state, amplified by 1 / beta.)
    tol = 1e-5 if fx.get("step_tol") is None and fx["serc"] <= 1e-8 else 1e-4
    for a, b in zip(tr, fx["trace"]):
        assert abs(a[1] - b[1]) < 1e-12 and not a[5]
        assert abs(a[3] / b[3] - 1.0) < tol, (a, b)           # relative power (north star: 1e-4)
        assert abs(a[2] - b[2]) < tol, (a, b)                 # reactivity [$]
    s.close()


def test_refined_lmw_fixture_sensitivity_is_below_the_bar():
    """the two committed oracle runs of the full-size LMW fixture (time steps converged to 1e-7 / 1e-6) bound the trace's
    dependence on the exit iteration: an order of magnitude below the 1e-4 bar of the GPU test"""
    import json
    from conftest import GOLDEN
    a = json.load(open(os.path.join(GOLDEN, "lmw_refined_full_oracle.json")))
    b = json.load(open(os.path.join(GOLDEN, "lmw_refined_full_oracle_t6.json")))
    assert a["serc"] == b["serc"] == 1e-8 and a["step_tol"] == 1e-7 and b["step_tol"] == 1e-6
    assert a["trace"][0][2] == b["trace"][0][2] and abs(a["trace"][0][2]) < 1e-5     # t = 0: same converged state, rho ~ 0
    for x, y in zip(a["trace"][1:], b["trace"][1:]):
        assert abs(x[3] / y[3] - 1.0) < 3e-5 and abs(x[2] - y[2]) < 3e-5


@pytest.mark.gpu
def test_gpu_c3_full_solve_against_cpu_oracle_fixture():
    """BASELINE configs[2] (IAEA-3D at 4 x 4 nodes per assembly, 190 planes: 183 160 nodes) with the reference's default
    nodal-update interval (nupd = 104) and nin = 4 against the CPU oracle's solve (tools/c3_oracle.py, one minute of
    CPU, committed): k-eff within 1 pcm, assembly power 1e-5, nodal power 1e-5.
    Both sides are converged to serc = ferc = 1e-8 (the fixture's "serc").  Round 1 compared at the 1e-5 exit of the oracle,
    at a fixed outer count; there the iterate still moves by 1.2e-5 per 20 iterations (nin = 2 sweeps are far from converged),
    so the two summation orders sat 1.4e-5 apart in assembly power -- a property of the unconverged iterate, not of the
    solution.  Converged, the trajectories contract onto the same solution and the north-star bars apply as they are."""
    import json
    from conftest import GOLDEN
    from adpres_b200 import capi
    ref = json.load(open(os.path.join(GOLDEN, "c3_oracle_result.json")))
    p = load_problem("IAEA3Ds").refine(xdiv=ref["xdiv"], ydiv=ref["ydiv"], zdiv=ref["zdiv"])
    assert (p.nnod, p.nupd) == (ref["nnod"], ref["nupd"]) and p.nin == 2 and ref["nin"] == 4
    assert ref["serc"] <= 1e-8 and ref["status"] == 0
    # nin = 4, twice the deck's default: with nin = 2 the two-node iteration is only marginally stable on this mesh (5 cm x 5 cm
    # x 2 cm nodes) -- the same solve takes 1 461 ... 2 573 outers depending on how the partial sums of the dot products are
    # grouped, with source-error excursions of 1e3 ... 1e5, and one order (NCCL path on two slabs) ran into the reference's own
    # "MAX. CHANGE > 1e3" STOP; from nin = 4 on all orders converge smoothly in 602 - 617 outers (tools/order_probe.py, round 2)
    # run for exactly the oracle's outer count (its 1e-8 exit, 612): the solution still moves by ~1e-5 in power from one
    # nodal update to the next (nupd = 104), so both sides must have seen the same number of updates
    s = capi.Solver(p, nout=ref["outers"], serc=0.0, ferc=0.0, nin=ref["nin"])
    rc, n = s.outer(0)
    assert rc == capi.STOP_MAXOUTER and n == ref["outers"], (rc, n, ref["outers"])
    st = s.state()
    assert st["ser"] < 10 * ref["serc"] and st["fer"] < 10 * ref["serc"]     # and it IS converged there
    assert abs(st["Ke"] - ref["keff"]) * 1e5 < 1.0
    rc, pw = s.powdis()
    asm, asm_ref = p.asm_power(pw), np.array(ref["asm_power"])
    nz = asm_ref > 0
    assert np.abs(asm[nz] / asm_ref[nz] - 1).max() < 1e-5
    idx = np.array(sorted(int(i) for i in ref["power_samples"]))
    ref_pw = np.array([ref["power_samples"][str(i)] for i in idx])
    nzp = ref_pw > 1e-12
    assert np.abs(pw[idx][nzp] / ref_pw[nzp] - 1).max() < 1e-5
    s.close()


@pytest.mark.gpu
def test_gpu_multigroup_adf_mid_size_against_cpu_oracle_fixture():
    """BASELINE configs[3] at a size where the nodal kernels for many groups matter: 8 groups with ADFs (tests/synth.py) on a
    5 cm mesh, 73 264 nodes = 586 k unknowns, about 220 000 two-node 16 x 16 systems per nodal update -- against the CPU oracle
    (tools/c4_mid_oracle.py, 80 s of CPU, committed).  Compared at the oracle's outer count, as for configs[2]."""
    import json
    import sys
    from conftest import GOLDEN, ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import iaea3d_multigroup
    from adpres_b200 import capi
    ref = json.load(open(os.path.join(GOLDEN, "c4_mid_oracle_result.json")))
    p = iaea3d_multigroup(ref["ng"]).refine(xdiv=ref["xdiv"], ydiv=ref["ydiv"], zdiv=ref["zdiv"])
    assert p.nnod == ref["nnod"]
    # fixed outer count (the oracle's 134; at its exit one iteration still moves the nodal power by 2.4e-5)
    s = capi.Solver(p, nin=ref["nin"], nupd=ref["nupd"], nac=ref["nac"], nout=ref["outers"], serc=0.0, ferc=0.0)
    rc, n = s.outer(0)
    assert ref["status"] == 0 and rc == capi.STOP_MAXOUTER and n == ref["outers"], (rc, n)
    assert abs(s.state()["Ke"] / ref["keff"] - 1.0) < 1e-5
    rc, pw = s.powdis()
    asm, asm_ref = p.asm_power(pw), np.array(ref["asm_power"])
    nz = asm_ref > 0
    assert np.abs(asm[nz] / asm_ref[nz] - 1).max() < 1e-5
    idx = np.array(sorted(int(i) for i in ref["power_samples"]))
    ref_pw = np.array([ref["power_samples"][str(i)] for i in idx])
    nzp = ref_pw > 1e-12
    assert np.abs(pw[idx][nzp] / ref_pw[nzp] - 1).max() < 5e-5
    s.close()


@pytest.mark.gpu
def test_gpu_multigroup_adf_full_size_against_cpu_oracle_fixture():
    """BASELINE configs[3] AT SIZE: 8 groups with ADFs (tests/synth.py) on the 1 cm x 1 cm x 2 cm mesh of configs[1] (4 579 000
    nodes, 36.6 M unknowns, 13.7 M two-node 16 x 16 systems per nodal update) against the CPU oracle (tools/c4_full_oracle.py,
    6 min of CPU, committed): four outer iterations from flat flux, ONE SANM nodal update, two more outer iterations on the
    updated matrix.  The fixture stays that short on purpose: from flat flux this iteration is far from contractive and
    amplifies the reduction-order round-off (k-eff GPU vs oracle 2e-9 at p = 1..4, 1e-6 after the extrapolation at p = 5, 1e-3
    at p = 45 -- measured with a 52-outer fixture), so later iterates are not comparable; here both sides still hold the same
    iterate and the nodal kernels see the same input.  Also: the quad kernels (default for G >= 5) and the
    one-thread-per-item kernels are bit-identical at this size."""
    import json
    import sys
    from conftest import GOLDEN, ROOT
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from synth import iaea3d_multigroup
    from adpres_b200 import capi
    ref = json.load(open(os.path.join(GOLDEN, "c4_full_oracle_result.json")))
    p = iaea3d_multigroup(ref["ng"]).refine(xdiv=ref["xdiv"], ydiv=ref["ydiv"], zdiv=ref["zdiv"])
    assert p.nnod == ref["nnod"] == 4579000
    nodes = np.array(ref["sample_nodes"])
    out = {}
    for form in (2, 0):
        ctl = dict(nin=ref["nin"], nupd=ref["nupd"], nac=ref["nac"], serc=0.0, ferc=0.0)
        s = capi.Solver(p, nout=ref["n1"], **ctl)
        s.set_option("nodal_coop", form)
        s.enable_trace()
        rc1, m1 = s.outer(0)
        rcn, nd, loc = s.nodal_upd(1)
        dn = s.nod()[1][:, nodes, :].copy()
        s.set_control(nout=ref["n2"], **ctl)
        rc2, m2 = s.outer(0)
        out[form] = dict(rc=[rc1, rcn, rc2], m=[m1, m2], nd=nd, loc=tuple(loc), ke=[r[1] for r in s.trace_rows], f0=s.state()["f0"], dn=dn)
        s.close()
    a, b = out[2], out[0]
    # the two kernel forms: same operations in the same order
    assert a["rc"] == b["rc"] and a["ke"] == b["ke"] and (a["nd"], a["loc"]) == (b["nd"], b["loc"])
    assert np.array_equal(a["dn"], b["dn"]) and np.array_equal(a["f0"], b["f0"])
    # against the oracle
    assert a["rc"] == ref["status"] and a["m"] == ref["outers"]
    kr = np.array(ref["trace_ke_first"] + ref["trace_ke"])
    assert np.abs(np.array(a["ke"]) / kr - 1).max() < C4_FULL_KE_TOL, (a["ke"], kr.tolist())
    assert a["loc"] == tuple(ref["ndloc"]) and abs(a["nd"] / ref["ndmax"] - 1) < C4_FULL_ND_TOL, (a["nd"], a["loc"], ref["ndmax"], ref["ndloc"])
    dnr = np.array(ref["dn_samples"])
    assert np.abs(a["dn"] - dnr).max() < C4_FULL_DN_TOL * np.abs(dnr).max()
    assert np.abs(a["f0"][nodes, :] / np.array(ref["f0_samples"]) - 1).max() < C4_FULL_F0_TOL


# bars of the full-size configs[3] comparison: ten times the differences measured on a B200 (tools/c4_full_gpu.py)
C4_FULL_KE_TOL, C4_FULL_ND_TOL, C4_FULL_DN_TOL, C4_FULL_F0_TOL = 1e-7, 1e-6, 1e-6, 1e-6
