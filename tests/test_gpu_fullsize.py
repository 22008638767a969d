"""GPU tests at BASELINE.json's full size (IAEA-3D refined to 1 cm x 1 cm x 2 cm, 4.579 M nodes,
9.158 M node-groups) through size-independent properties, plus direct oracle comparisons on the
bounded sample the CPU can finish in seconds (same radial mesh, 19 planes, 457 900 nodes)."""
import numpy as np
import pytest

from conftest import load_problem

pytestmark = pytest.mark.gpu


def _refined(zdiv):
    return load_problem("IAEA3Ds").refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=zdiv)


@pytest.fixture(scope="module")
def c2():
    from adpres_b200 import capi
    p = _refined([10] * 19)
    assert (p.nxx, p.nyy, p.nzz, p.nnod) == (170, 170, 190, 4579000)
    s = capi.Solver(p, nin=2, nac=5, nupd=50, nout=100000)
    s.matrix_setup(1)
    return p, s


def _numpy_spmv(p, a, x, g):
    """7-diagonal SpMV in numpy from the matrix the device assembled (set_ind order)."""
    npl, N = p.npl, p.nnod
    i, j = p.ix[:npl], p.iy[:npl]
    nodp = np.zeros((p.nxx + 2, p.nyy + 2), dtype=np.int64)
    nodp[i, j] = np.arange(1, npl + 1)
    ypm = np.where(j == p.xstag_smin[i - 1], 0, np.arange(1, npl + 1) - nodp[i, j - 1])
    ypp = np.where(j == p.xstag_smax[i - 1], 0, nodp[i, j + 1] - np.arange(1, npl + 1))
    xp = np.concatenate([np.zeros(npl), x, np.zeros(npl)])
    idx = np.arange(N) + npl
    r = np.arange(N) % npl
    y = np.zeros(N)
    for d, off in enumerate((-npl, -ypm[r], -1, 0, 1, ypp[r], npl)):
        y = y + a[d, :, g] * xp[idx + off]
    return y


def test_c2_spmv_bit_exact_against_numpy_and_linear(c2):
    p, s = c2
    a = s.matrix_dia()
    rng = np.random.default_rng(5)
    x, y = rng.standard_normal(p.nnod), rng.standard_normal(p.nnod)
    for g in (1, 2):
        ax = s.sp_matvec(g, x)
        assert np.array_equal(ax, _numpy_spmv(p, a, x, g - 1))
        ay = s.sp_matvec(g, y)
        lin = s.sp_matvec(g, 2.0 * x - 0.5 * y)
        assert np.abs(lin - (2.0 * ax - 0.5 * ay)).max() <= 1e-12 * np.abs(ax).max()


def test_c2_matrix_structure(c2):
    """FDM matrix (dn = 0): positive diagonal, non-positive off-diagonals, weak diagonal dominance
    with margin sigr, and symmetry of the coupling df(n,+) = df(p,-) across every face."""
    p, s = c2
    a = s.matrix_dia()
    df, dn = s.nod()
    assert not dn.any()
    for g in range(p.ng):
        diag = a[3, :, g]
        off = np.delete(a[:, :, g], 3, axis=0)
        assert (diag > 0).all() and (off <= 0).all()
        assert (diag + off.sum(axis=0) >= p.sigr[:, g] * (1 - 1e-12)).all()
    npl = p.npl
    assert np.array_equal(df[4, :-npl, :], df[5, npl:, :])          # z faces
    xin = p.ix[:-1] != p.ystag_smax[p.iy[:-1] - 1]
    assert np.array_equal(df[0, :-1, :][xin], df[1, 1:, :][xin])    # x faces


def test_c2_bicg_reduces_residual(c2):
    p, s = c2
    rng = np.random.default_rng(9)
    b = rng.random(p.nnod)
    x0 = np.zeros(p.nnod)
    r0 = np.linalg.norm(b)
    prev = r0
    for imax in (2, 8):
        x = s.bicg(imax, 2, b, x0)
        res = np.linalg.norm(b - s.sp_matvec(2, x))
        assert res < prev
        prev = res
    assert prev < 0.5 * r0


def test_c2_outer_iterations_power_and_symmetry(c2):
    """30 outer iterations at full size: k-eff finite and in range, power normalised, and the
    IAEA-3D core's diagonal symmetry (i,j) -> (171-j,171-i) preserved."""
    p, s = c2
    s.init_flux()
    s.outer_begin(0)
    for q in range(1, 31):
        ke, ser, fer = s.outer_iter(0, q)
    assert 0.9 < ke < 1.1 and np.isfinite(ser) and np.isfinite(fer)
    rc, pw = s.powdis()
    assert rc == 0 and abs(pw.sum() - 1.0) < 1e-10 and (pw >= 0).all()
    fx = np.zeros((p.nxx + 1, p.nyy + 1, p.nzz + 1))
    fx[p.ix, p.iy, p.iz] = pw
    mirror = fx[171 - p.iy, 171 - p.ix, p.iz]
    assert np.abs(pw - mirror).max() < 1e-6 * pw.max()
    # device-resident stepping gives the same iterates as the per-iteration API
    from adpres_b200 import capi
    s2 = capi.Solver(p, nin=2, nac=5, nupd=50, nout=100000)
    s2.matrix_setup(1); s2.init_flux(); s2.outer_begin(0)
    rc, ke2, ser2, fer2 = s2.outer_steps(0, 1, 30)
    assert (ke2, ser2, fer2) == (ke, ser, fer)


def test_sample_mesh_iterations_and_nodal_update_vs_oracle():
    """Same radial mesh, 19 planes (457 900 nodes): 12 outer iterations and one SANM nodal update
    on GPU and oracle from the same start."""
    from adpres_b200 import capi
    from oracle import Oracle
    p = _refined([1] * 19)
    kw = dict(nin=2, nac=5, nupd=10, nout=12)
    s, o = capi.Solver(p, **kw), Oracle(p, **kw)
    s.enable_trace()
    rc_s, n_s = s.outer(1)
    rc_o, n_o = o.outer(1)
    assert n_s == n_o == 12
    ke_o, ser_o, fer_o = o.trace()
    for (q, ke, ser, fer) in s.trace_rows:
        assert abs(ke - ke_o[q - 1]) < 1e-8
        # the maxima sit where the new value is ~0 (|new-old|/|new| up to 1e3+): noise there is amplified
        assert abs(ser / ser_o[q - 1] - 1) < (1e-6 if ser_o[q - 1] < 1 else 1e-3)
        assert abs(fer / fer_o[q - 1] - 1) < (1e-6 if fer_o[q - 1] < 1 else 1e-3)
    # At 1 cm nodes the SANM constants B, E, G are ill-conditioned in fp64 (alpha ~ 0.07-0.23:
    # terms of size 3/alpha^3 cancel to O(alpha^4); SURVEY.md section 7 measured a 1-ulp change of
    # sinh moving them by up to 5e-6).  The device libm's sinh/cosh differ from glibc's in the
    # last bit, so dn agrees to ~1e-7 here instead of the 1e-9 seen on 10-20 cm nodes.
    assert abs(s.trace_nodal[0][1] / o.nodal_trace()[0][1] - 1) < 1e-6
    dn_s, dn_o = s.nod()[1], o.nod()[1]
    assert np.abs(dn_s - dn_o).max() < 1e-6
    f_s, f_o = s.state()["f0"], o.state()["f0"]
    assert np.abs(f_s - f_o).max() / np.abs(f_o).max() < 1e-7


@pytest.mark.parametrize("fixture", ["c2_oracle_result.json", "c2prime_oracle_result.json"])
def test_c2_full_solve_against_cpu_oracle_fixture(fixture):
    """BASELINE.json configs[1] at full size (C2: 4.58 M nodes, 2 cm planes) and the north star's
    ">= 10 M nodes" variant (C2': 10.07 M nodes, 0.91 cm planes) against the CPU oracle.  The
    oracle needs 45 / 63 min of CPU for these solves, so its results are committed fixtures
    (tests/golden/c2*_oracle_result.json, made by tools/oracle_fullsize.py in the container that
    has the oracle and the time); the GPU runs the identical %ITER card.  North-star bars: k-eff
    within 1 pcm, power within 1e-5."""
    import json
    import os
    from conftest import GOLDEN
    from adpres_b200 import capi
    path = os.path.join(GOLDEN, fixture)
    if not os.path.exists(path):
        pytest.skip("full-size oracle fixture not generated")
    ref = json.load(open(path))
    p = _refined([ref["zdiv"]] * 19)
    assert p.nnod == ref["nnod"]
    # %ITER of the fixture: nin = 10 on the 2 cm planes; nin = 20 on the 0.91 cm planes of C2' -- there nin = 10 sits on the
    # edge of the two-node iteration's stability (source-error excursions of 1e4 - 1e7 in every summation order tried, and a
    # "MAX. CHANGE > 1e3" STOP in one of them: tools/c2prime_probe.py), nin = 15 / 20 converge smoothly in all of them
    s = capi.Solver(p, nin=ref.get("nin", 10), nac=5, nupd=50, nout=3000)
    s.enable_trace()
    rc, n = s.outer(1)
    assert rc == ref["status"] == 0
    ke = s.state()["Ke"]
    assert abs(ke - ref["keff"]) * 1e5 < 1.0, (ke, ref["keff"])
    # Iteration path: ten unconverged BiCGSTAB sweeps per outer amplify reduction-order round-off
    # (measured: |dKe| 1e-10 at p = 1-3, 1e-8 at p = 5, 1e-6 at p = 20, 1e-4 at p = 50) before both
    # runs contract onto the same solution (378 vs 383 outers, k-eff equal to 2e-9).
    if ref["zdiv"] == 10:
        for (q, k, ser, fer) in s.trace_rows[:15]:
            assert abs(k - ref["trace_ke"][q - 1]) < (1e-9 if q <= 3 else 1e-6), (q, k, ref["trace_ke"][q - 1])
        # the first nodal update (p = 50) still sees nearly the same iterate; by the second (p = 100) the two summation
        # orders have drifted apart (round 2: |d ndmax| 80 % and another location after the tile order of three kernels
        # changed, with k-eff and power at convergence unchanged) -- only the first is a parity check
        mine, theirs = s.trace_nodal[0], ref["nodal_updates"][0]
        assert mine[0] == theirs[0] and abs(mine[1] / theirs[1] - 1) < 1e-2, (mine, theirs)
        # the outer COUNT is not a parity quantity on this mesh: 378 in the serial oracle, 383 with round 1's tile order,
        # 274 with round 2's (boundary-plane tiles first in three kernels) -- the exit falls into one or another nodal-update
        # cycle (nupd = 50) depending on round-off, while k-eff and the power at the exit agree to 1e-9 / 5e-6
        # (202 with the reversed sweeps of B and D + the grouped-load C kernel)
        assert 100 <= n <= 800, n
    else:
        # 0.91 cm planes, nin = 20: the twenty BiCGSTAB sweeps of an outer iteration run past the point where the inner
        # solve has converged; BiCGSTAB without a residual test (mod_cmfd.f90:1203-1243) then amplifies the reduction-order
        # round-off inside ONE outer iteration (|dKe| 1.5e-6 at p = 1, where the nin = 10 card had 3e-9) while the outer
        # iteration itself contracts: same outer count (156 - 159 in every order, oracle 159), k-eff within 0.002 pcm
        for (q, k, ser, fer) in s.trace_rows[:10]:
            assert abs(k - ref["trace_ke"][q - 1]) < 1e-4, (q, k, ref["trace_ke"][q - 1])
        assert abs(n - ref["outers"]) <= 0.1 * ref["outers"], (n, ref["outers"])
        mine, theirs = s.trace_nodal[0], ref["nodal_updates"][0]
        assert mine[0] == theirs[0] and abs(mine[1] / theirs[1] - 1) < 1e-2, (mine, theirs)
    rc, pw = s.powdis()
    asm, asm_ref = p.asm_power(pw), np.array(ref["asm_power"])
    nz = asm_ref > 0
    assert np.abs(asm[nz] / asm_ref[nz] - 1).max() < 1e-5
    idx = np.array(sorted(int(i) for i in ref["power_samples"]))
    ref_pw = np.array([ref["power_samples"][str(i)] for i in idx])
    nzp = ref_pw > 1e-12
    assert np.abs(pw[idx][nzp] / ref_pw[nzp] - 1).max() < 1e-5     # north star: nodal power within 1e-5 (measured 5e-6)
