"""Worker for the multi-GPU z-slab tests (launched by torchrun, one rank per GPU).
Every rank holds the global problem (as N copies of the Fortran driver would), owns a z-slab,
and checks its own slab against the single-process CPU oracle."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    from conftest import load_problem
    from adpres_b200 import capi
    from oracle import Oracle
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = (capi.C.c_ubyte * 128)()
    if rank == 0:
        assert capi.load().adp_comm_unique_id(buf) == 0
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    uid = bytes(t.cpu().tolist())

    deck = sys.argv[1] if len(sys.argv) > 1 else "IAEA3Ds"
    if deck == "LMW_tr":
        return transient_main(rank, world, local, uid, dist)
    if deck == "NEACRP_th":
        return th_main(rank, world, local, uid, dist)
    if deck == "NEACRP_cb":
        return cb_main(rank, world, local, uid, dist)
    if deck == "MOX_xtab":
        return xtab_main(rank, world, local, uid, dist)
    if deck == "C3_fixture":
        return c3_main(rank, world, local, uid, dist)
    if deck == "SYNTH8_adf":
        return multigroup_main(rank, world, local, uid, dist)
    if deck == "LMW_refined":
        return lmw_refined_main(rank, world, local, uid, dist)
    if deck == "IAEA3Ds_z2":                  # 38 planes: uneven slabs at 4 ranks, 2 planes per axial assembly
        p = load_problem("IAEA3Ds").refine(zdiv=[2] * 19)
    else:
        p = load_problem(deck)
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid)
    o = Oracle(p)
    own = s.own
    rel = lambda a, b: np.abs(a - b).max() / max(np.abs(b).max(), 1e-300)

    # 1. matrix + SpMV on the owned rows: bit exact
    s.matrix_setup(1); o.matrix_setup(1)
    assert np.array_equal(s.matrix_dia()[:, own, :], o.matrix_dia()[:, own, :])
    x = np.random.default_rng(3).standard_normal(p.nnod)
    for g in range(1, p.ng + 1):
        assert np.array_equal(s.sp_matvec(g, x)[own], o.sp_matvec(g, x)[own]), "spmv halo"
    # 2. BiCGSTAB with halo exchange + all-reduced dot products
    b = np.random.default_rng(4).random(p.nnod); x0 = np.random.default_rng(5).random(p.nnod)
    assert rel(s.bicg(3, 1, b, x0)[own], o.bicg(3, 1, b, x0)[own]) < 1e-11
    # 3. one nodal update from an identical state (interface surfaces computed redundantly)
    o.set_control(nout=6, nupd=1000); o.outer(0)
    st = o.state()
    s.set_state(st["f0"], st["fs0"], st["Ke"])
    rc_o = o.nodal_upd(1)
    rc_s, ndmax, loc = s.nodal_upd(1)
    assert rc_o == 0 and rc_s == 0
    dn_s, dn_o = s.nod()[1], o.nod()[1]
    assert np.abs(dn_s[:, own, :] - dn_o[:, own, :]).max() < 1e-9, np.abs(dn_s[:, own, :] - dn_o[:, own, :]).max()
    assert abs(ndmax - o.ndmax) < 1e-9 * max(1.0, o.ndmax), (ndmax, o.ndmax)
    # 4. the whole eigenvalue solve
    del s
    dist.barrier()
    buf2 = (capi.C.c_ubyte * 128)()
    if rank == 0:
        assert capi.load().adp_comm_unique_id(buf2) == 0
    t = torch.tensor(list(bytes(buf2)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=bytes(t.cpu().tolist()))
    o = Oracle(p)
    rc_s, n_s = s.outer(0)
    rc_o, n_o = o.outer(0)
    assert rc_s == rc_o == 0 and abs(n_s - n_o) <= 1, (rc_s, rc_o, n_s, n_o)
    ks, ko = s.state()["Ke"], o.state()["Ke"]
    assert abs(ks - ko) * 1e5 < 1.0, (ks, ko)
    _, pw_s = s.powdis()
    _, pw_o = o.powdis()
    nz = pw_o[own] > 1e-12
    assert np.abs(pw_s[own][nz] / pw_o[own][nz] - 1).max() < 1e-5
    # 5. result reductions (AsmPow / AxiPow / AsmFlux): the slabs' partial column and plane sums are
    #    all-reduced, every rank gets the whole map.  Inputs for the oracle: the GPU's own power and
    #    flux, gathered from the slabs.
    from oracle import results

    def gather(a):
        full = np.zeros_like(a)
        full[own] = a[own]
        tt = torch.from_numpy(np.ascontiguousarray(full.T)).cuda()
        dist.all_reduce(tt)
        return np.asfortranarray(tt.cpu().numpy().T) if a.ndim == 2 else tt.cpu().numpy()
    pw_full, f0_full = gather(pw_s), gather(s.state()["f0"])
    fasm, im, jm = s.asm_pow()
    ref, ri, rj = results.asm_pow(p, pw_full)
    assert np.abs(fasm - ref).max() < 1e-13 and (im, jm) == (ri, rj)
    faxi, am = s.axi_pow()
    ra, ram = results.axi_pow(p, pw_full)
    assert np.abs(faxi - ra).max() < 1e-13 and am == ram
    fa, neg = s.asm_flux()
    rf, rneg = results.asm_flux(p, f0_full)
    assert np.abs(fa - rf).max() < 1e-13 * np.abs(rf).max() and neg == rneg
    print(f"RANK {rank}/{world} OK deck={deck} planes=[{s.k0},{s.k1}) keff={ks:.6f} outers={n_s}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def transient_main(rank, world, local, uid, dist):
    """Device-resident rod-ejection time steps (XS update, %EXTR, time-step glue, outer_tr, uPden,
    PowTot, reactivity with Lxyz) on z-slabs against the single-process oracle with numpy glue."""
    from conftest import load_problem
    from adpres_b200 import capi, transient
    from oracle import Oracle

    def tight(p):
        p.serc = p.ferc = 1e-9
        p.nout = 5000
        p.bextr = 1
        return p
    p1, p2 = tight(load_problem("LMW")), tight(load_problem("LMW"))
    tr_o = transient.rod_eject(p1, Oracle(p1), max_steps=4)
    s = capi.Solver(p2, device=local, nranks=world, rank=rank, uid=uid)
    tr_d = transient.rod_eject_device_glue(p2, s, max_steps=4, device_xs=True)
    for a, b in zip(tr_d, tr_o):
        assert a[1] == b[1]
        assert abs(a[3] / b[3] - 1) < 1e-5, (a, b)
        assert abs(a[2] - b[2]) < 1e-5, (a, b)
    print(f"RANK {rank}/{world} OK deck=LMW_tr planes=[{s.k0},{s.k1}) power={tr_d[-1][3]:.6f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def lmw_refined_main(rank, world, local, uid, dist):
    """BASELINE configs[4] on z-slabs: the LMW rod ejection refined to 2 cm x 2 cm x 4 cm (146 250 nodes, nin = 10, nupd = 50),
    every solve converged to 1e-8, device-resident time stepping, against the committed CPU-oracle trace
    (tests/golden/lmw_refined_mid_oracle.json, tools/lmw_refined_oracle.py)."""
    import json
    from conftest import GOLDEN, load_problem
    from adpres_b200 import capi, transient
    fx = json.load(open(os.path.join(GOLDEN, "lmw_refined_mid_oracle.json")))
    rdiv, zdiv = fx["rdiv"], fx["zdiv"]
    p = load_problem("LMW").refine(xdiv=[rdiv // 2] + [rdiv] * 5, ydiv=[rdiv // 2] + [rdiv] * 5, zdiv=[zdiv] * 10)
    p.nin, p.nupd, p.nac, p.nout, p.serc, p.ferc, p.biter = 10, 50, 5, 20000, fx["serc"], fx["serc"], 1
    assert p.nnod == fx["nnod"]
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid)
    tr = transient.rod_eject_device_glue(p, s, max_steps=len(fx["trace"]) - 1, device_xs=True, step_tol=fx.get("step_tol"))
    assert len(tr) == len(fx["trace"])
    for a, b in zip(tr, fx["trace"]):
        assert abs(a[1] - b[1]) < 1e-12 and not a[5]
        assert abs(a[3] / b[3] - 1.0) < 1e-5, (a, b)           # relative power (north star: 1e-4)
        assert abs(a[2] - b[2]) < 1e-4, (a, b)                 # reactivity in dollars
    print(f"RANK {rank}/{world} OK deck=LMW_refined planes=[{s.k0},{s.k1}) power={tr[-1][3]:.6f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def th_main(rank, world, local, uid, dist):
    """th_upd / th_trans on z-slabs: the enthalpy march is a chain, every slab receives entm / bfrate
    from the slab below and passes them on.  Owned nodes against the single-process oracle."""
    import numpy as np
    from adpres_b200 import capi
    from oracle import th as oth
    from test_th import _setup
    p, th, st, npow = _setup()
    xpl = oth.pline_static(p, th, npow)
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid)
    s.set_th(th)
    s.set_th_state(st)
    own = s.own

    def check():
        g = s.th_state()
        for k in ("tfm", "heatf", "ent", "ftem", "mtem", "cden", "frate"):
            if st.get(k) is None:
                continue
            d = np.abs(g[k][own] - st[k][own]).max() / max(np.abs(st[k]).max(), 1e-300)
            assert d < 1e-12, (k, d)
    for it in range(4):
        old = st["ftem"].copy()
        oth.th_upd(p, th, st, xpl)
        rc, err = s.th_upd(xpl)
        assert rc == 0 and abs(err - oth.abs_e(st["ftem"], old)) < 1e-9 * max(1.0, err)      # all-reduced maximum
        check()
    for it in range(3):
        oth.th_trans(p, th, st, xpl * 1.4, 0.05)
        assert s.th_trans(xpl * 1.4, 0.05) == 0
        check()
    print(f"RANK {rank}/{world} OK deck=NEACRP_th planes=[{s.k0},{s.k1}) ftem_max={st['ftem'].max():.3f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def cb_main(rank, world, local, uid, dist):
    """Critical boron search with TH feedback, everything device-resident, on z-slabs: the feedback XS
    update needs the neighbours' temperatures / densities on the ghost planes, the TH march is a chain,
    pline comes from the all-reduced PowDis.  Case C2 (half core, full power, partially inserted rods)
    against the reference's own value and the single-process oracle."""
    import json
    import numpy as np
    from conftest import GOLDEN, load_problem
    from adpres_b200 import capi, thermal
    from oracle import Oracle, th as oth
    gold = json.load(open(os.path.join(GOLDEN, "neacrp_bcon.json")))["ppm"]["C2"]
    p1, p2 = load_problem("NEACRP_C2"), load_problem("NEACRP_C2")
    s = capi.Solver(p2, device=local, nranks=world, rank=rank, uid=uid)
    gd = thermal.DeviceGlue(p2, s)
    bd, rd = thermal.cbsearcht(gd)
    assert abs(bd - gold) < 0.1, (bd, gold)
    go = thermal.HostGlue(p1, Oracle(p1), oth)
    bo, ro = thermal.cbsearcht(go)
    assert len(ro) == len(rd) and abs(bo - bd) < 0.02, (bo, bd)
    fo, fd = go.th_fields(), gd.th_fields()
    own = s.own
    # full-power case: th_iter leaves its loop where th_err crosses 0.01 K, which may be one pass earlier or later for
    # another summation order (tests/test_th.py uses the same 1e-4 for its power case on one GPU)
    for k in fo:
        dev = np.abs(fd[k][own] / fo[k][own] - 1.0).max()
        assert dev < 1e-4, (k, dev)
    print(f"RANK {rank}/{world} OK deck=NEACRP_cb planes=[{s.k0},{s.k1}) bcon={bd:.2f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def multigroup_main(rank, world, local, uid, dist):
    """BASELINE configs[3] on z-slabs: the synthetic 8-group deck with ADFs (tests/synth.py; 2 planes per axial assembly so that
    4 and 8 ranks get slabs too) -- SpMV bit exact on the owned rows, one SANM nodal update from an identical state (the quad
    kernels of round 2 incl. the surface shared by two slabs, computed redundantly on both ranks), and twelve outer iterations
    with three nodal updates at a fixed count against the single-process oracle."""
    from adpres_b200 import capi
    from oracle import Oracle
    from synth import iaea3d_multigroup
    p = iaea3d_multigroup(8, zdiv=[2] * 19)
    ctl = dict(nin=4, nupd=4, nac=1000, nout=12, serc=0.0, ferc=0.0)     # (no source extrapolation: it amplifies round-off)
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid, **ctl)
    o = Oracle(p, **ctl)
    own = s.own
    s.matrix_setup(1); o.matrix_setup(1)
    assert np.array_equal(s.matrix_dia()[:, own, :], o.matrix_dia()[:, own, :])
    x = np.random.default_rng(3).standard_normal(p.nnod)
    for g in (1, 4, 8):
        assert np.array_equal(s.sp_matvec(g, x)[own], o.sp_matvec(g, x)[own]), "spmv halo"
    # one nodal update from an identical state
    o.set_control(nout=3, nupd=1000, nin=4, nac=1000, serc=0.0, ferc=0.0); o.outer(0)
    st = o.state()
    s.set_state(st["f0"], st["fs0"], st["Ke"])
    rc_o = o.nodal_upd(1)
    rc_s, ndmax, loc = s.nodal_upd(1)
    assert rc_o == 0 and rc_s == 0
    dn_s, dn_o = s.nod()[1], o.nod()[1]
    assert np.abs(dn_s[:, own, :] - dn_o[:, own, :]).max() < 1e-9, np.abs(dn_s[:, own, :] - dn_o[:, own, :]).max()
    assert abs(ndmax - o.ndmax) < 1e-9 * max(1.0, o.ndmax) and tuple(loc) == tuple(o.ndloc()), (ndmax, o.ndmax, loc, o.ndloc())
    # fixed-count run with three nodal updates
    import torch
    del s
    dist.barrier()
    buf2 = (capi.C.c_ubyte * 128)()
    if rank == 0:
        assert capi.load().adp_comm_unique_id(buf2) == 0
    t = torch.tensor(list(bytes(buf2)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=bytes(t.cpu().tolist()), **ctl)
    s.enable_trace()
    o = Oracle(p, **ctl)
    rc_s, n_s = s.outer(1)          # popt = 1: the nodal-update line is "printed" (traced) as well
    rc_o, n_o = o.outer(0)
    assert rc_s == rc_o == capi.STOP_MAXOUTER and n_s == n_o == 12, (rc_s, rc_o, n_s, n_o)
    ke_s, ke_o = np.array([r[1] for r in s.trace_rows]), o.trace()[0]
    assert np.abs(ke_s / ke_o - 1).max() < 1e-7, np.abs(ke_s / ke_o - 1).max()
    nt_s, nt_o = s.trace_nodal, o.nodal_trace()
    assert len(nt_s) == len(nt_o) == 3
    for a, b in zip(nt_s, nt_o):
        assert a[0] == b[0] and tuple(a[2:]) == tuple(b[2:]) and abs(a[1] / b[1] - 1) < 1e-6, (a, b)
    f_s, f_o = s.state()["f0"], o.state()["f0"]
    assert np.abs(f_s[own] / f_o[own] - 1).max() < 1e-6
    dn_s, dn_o = s.nod()[1], o.nod()[1]
    assert np.abs(dn_s[:, own, :] - dn_o[:, own, :]).max() < 1e-7
    print(f"RANK {rank}/{world} OK deck=SYNTH8_adf planes=[{s.k0},{s.k1}) keff={ke_s[-1]:.6f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def c3_main(rank, world, local, uid, dist):
    """BASELINE configs[2] -- IAEA-3D at 4 x 4 nodes per assembly, 190 planes, "z-slab over 2/4/8 B200" -- against the
    committed CPU-oracle solve (tests/golden/c3_oracle_result.json, converged to 1e-8)."""
    import json
    import numpy as np
    from conftest import GOLDEN, load_problem
    from adpres_b200 import capi
    ref = json.load(open(os.path.join(GOLDEN, "c3_oracle_result.json")))
    p = load_problem("IAEA3Ds").refine(xdiv=ref["xdiv"], ydiv=ref["ydiv"], zdiv=ref["zdiv"])
    # the oracle's converged solve (serc = ferc = 1e-8: 612 outers) repeated for EXACTLY its outer count.  nin of the
    # fixture (4): with the deck's default nin = 2 the two-node iteration is only marginally stable on this mesh (1 461 ...
    # 2 573 outers and, in one summation order, the reference's own ndmax > 1e3 STOP: tools/order_probe.py).  Fixed count:
    # the solution still moves by ~1e-5 in power from one nodal update to the next (nupd = 104), so both sides must have
    # seen the same number of updates -- two slabs meet the 1e-8 exit test one update cycle later (702 outers) and sit
    # 1.3e-5 away in assembly power without being any less converged
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid, nout=ref["outers"], serc=0.0, ferc=0.0, nin=ref["nin"])
    rc, n = s.outer(0)
    assert rc == capi.STOP_MAXOUTER and n == ref["outers"], (rc, n, ref["outers"])
    ke = s.state()["Ke"]
    assert abs(ke - ref["keff"]) * 1e5 < 1.0, (ke, ref["keff"])
    _, pw = s.powdis()
    lo, hi = s.own.start, s.own.stop
    idx = np.array(sorted(int(i) for i in ref["power_samples"] if lo <= int(i) < hi))
    ref_pw = np.array([ref["power_samples"][str(i)] for i in idx])
    nz = ref_pw > 1e-12
    assert np.abs(pw[idx][nz] / ref_pw[nz] - 1).max() < 1e-5
    fasm, _, _ = s.asm_pow()                      # all-reduced over the slabs: every rank holds the whole map
    asm_ref = np.array(ref["asm_power"])
    nzm = asm_ref > 0
    assert np.abs(fasm[nzm] / asm_ref[nzm] - 1).max() < 1e-5
    print(f"RANK {rank}/{world} OK deck=C3_fixture planes=[{s.k0},{s.k1}) keff={ke:.6f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


def xtab_main(rank, world, local, uid, dist):
    """MOX part 3 (%XTAB branch tables, rods from the rodded sets, TH feedback) on z-slabs: the table
    interpolation on the ghost planes takes the neighbours' fuel temperature / coolant density
    (adp_comm_halo inside adp_xs_update_xtab); a STOP on one rank must come back on all.  Against the
    reference's own critical boron (1341.99 ppm)."""
    import json
    import numpy as np
    from conftest import GOLDEN, load_problem
    from adpres_b200 import capi, thermal
    gold = json.load(open(os.path.join(GOLDEN, "mox_bcon.json")))["ppm"]["P3_HELIOS"]
    p = load_problem("MOX_P3_HELIOS")
    s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid)
    gd = thermal.DeviceGlue(p, s)
    bd, rd = thermal.cbsearcht(gd)
    assert abs(bd - gold) < 0.02, (bd, gold)
    # out-of-range coolant density in ONE node of the LAST rank's slab: every rank must report the reference's STOP.
    # (The node must belong to a material whose table has a density branch -- the top planes are reflector with a
    # single-point table, where brInterp has nothing to range-check: round 1 put the bad value there and, never having
    # run on two GPUs, did not notice that no STOP can come from it.)
    n = p.nnod
    cden = np.full(n, 0.7)
    branched = np.array([t["nd"] > 1 for t in p.xtab])[p.mat - 1]
    bad = int(np.nonzero(branched)[0].max())
    assert bad >= (p.nzz - p.nzz // world) * p.npl - p.npl, "the node must lie in the last rank's slab"
    cden[bad] = 0.3
    rc = s.xs_update_xtab(bd, np.full(n, 900.0), np.full(n, 560.0), cden, p.crod["bpos"].astype(np.float64))
    assert rc == capi.STOP_XTAB_RANGE, rc
    print(f"RANK {rank}/{world} OK deck=MOX_xtab planes=[{s.k0},{s.k1}) bcon={bd:.2f}", flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
