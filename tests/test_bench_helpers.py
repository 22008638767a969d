"""bench.py's weak-scaling workload: rank r of N builds only its z-slab of the N-fold axially stacked core
(SlabProblem) instead of the whole global arrays.  Checked here against the explicitly stacked problem on a
small base mesh (CPU)."""
import dataclasses
import importlib.util
import os

import numpy as np
import pytest

from conftest import ROOT, load_problem


@pytest.fixture(scope="module")
def bench():
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    m = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(m)
    return m


@pytest.mark.parametrize("world", [2, 3, 8])
def test_slab_problem_equals_the_stacked_core(bench, world):
    from adpres_b200.slab import slab_planes
    base = load_problem("IAEA3Ds").refine(xdiv=[1] + [2] * 8, ydiv=[2] * 8 + [1], zdiv=[2] * 19)
    base1 = load_problem("IAEA3Ds").refine(xdiv=[1] + [2] * 8, ydiv=[2] * 8 + [1], zdiv=[1] * 19)
    p0 = load_problem("IAEA3Ds")
    full = dataclasses.replace(p0, nz=p0.nz * world, zsize=np.tile(p0.zsize, world), zdiv=np.tile(p0.zdiv, world),
                               zpln=np.tile(p0.zpln, world)).refine(xdiv=[1] + [2] * 8, ydiv=[2] * 8 + [1], zdiv=[2] * 19 * world)
    assert full.nnod == base.nnod * world and full.nzz == base.nzz * world
    covered = np.zeros(full.nnod, dtype=int)
    for rank in range(world):
        sp = bench.SlabProblem(base, world, rank)
        # the same slab from the one-plane-per-assembly base (what bench.py uses: zrefine planes per base plane)
        sp1 = bench.SlabProblem(base1, world, rank, stack=world, zrefine=2)
        for k in ("nnod", "nzz", "k0", "k1", "rows"):
            assert getattr(sp1, k) == getattr(sp, k), k
        assert np.array_equal(sp1.zdel, sp.zdel) and np.array_equal(sp1.mat[sp.rows], sp.mat[sp.rows])
        for k in ("D", "sigr", "nuf", "sigf", "exsrc", "sigs", "dc"):
            assert np.array_equal(getattr(sp1, k)[sp.rows], getattr(sp, k)[sp.rows]), k
        assert (sp.nnod, sp.nzz, sp.npl, sp.ng) == (full.nnod, full.nzz, full.npl, full.ng)
        assert np.array_equal(sp.zdel, full.zdel) and np.array_equal(sp.ix, full.ix) and np.array_equal(sp.iy, full.iy)
        assert np.array_equal(sp.iz, full.iz)
        k0, k1 = slab_planes(full.nzz, world, rank)
        assert (sp.k0, sp.k1) == (k0, k1)
        rows = sp.rows
        # own planes plus two ghost planes per side (what the library uploads)
        assert rows.start == max(0, k0 - 2) * full.npl and rows.stop == min(full.nzz, k1 + 2) * full.npl
        assert np.array_equal(sp.mat[rows], full.mat[rows])
        for k in ("D", "sigr", "nuf", "sigf", "exsrc", "sigs", "dc"):
            assert np.array_equal(getattr(sp, k)[rows], getattr(full, k)[rows]), k
        covered[k0 * full.npl:k1 * full.npl] += 1
    assert np.all(covered == 1)                       # the slabs tile the stacked core exactly once


def test_load_c2_sample_is_the_bounded_cpu_workload(bench):
    p = bench.load_c2(sample_planes=19)
    assert (p.nxx, p.nyy, p.nzz, p.nnod) == (170, 170, 19, 457900)
    assert np.all(p.xdel == 1.0) and np.all(p.zdel == 20.0)
    assert bench.CTL["nin"] == 10 and bench.CTL["nupd"] == 50 and bench.SPMV_BYTES_PER_ROW == 72.0


def test_timed_window_contains_a_nodal_update(bench):
    # driver call: --steps 20 --warmup 5, nupd = 50 -> timed iterations 31..50, one nodal update inside
    assert bench.window(20, 5, 50) == 31 and bench.updates_in(31, 20, 50) == 1
    assert bench.window(200, 5, 50) == 6 and bench.updates_in(6, 200, 50) == 4      # long runs: right after the warm-up
    assert bench.window(45, 5, 50) == 6 and bench.updates_in(6, 45, 50) == 1
    for K in (1, 3, 20, 44, 45, 50, 100):
        for W in (0, 3, 5, 10):
            P0 = bench.window(K, W, 50)
            assert P0 >= W + 1 and bench.updates_in(P0, K, 50) >= 1
    # both arms print the same config dict
    assert bench.workload_config(1, 20, 5) == bench.workload_config(1, 20, 5)
    assert bench.workload_config(1, 20, 5)["timed_iterations"] == [31, 50]


def test_clock_sampler_parses_nvidia_smi_lines(bench):
    s = bench.ClockSampler(0)
    s.proc = type("P", (), {"terminate": lambda self: None, "wait": lambda self, timeout=None: 0, "kill": lambda self: None})()
    s.lines = ["0, 1965, 1965, 1003.50, Not Active, Not Active, Not Active, Active",
               "0, 1950, 1965, 998.10, Not Active, Not Active, Not Active, Not Active",
               "0, [N/A], 1965, 10.0, Not Active, Not Active, Not Active, Not Active",       # unparsable sample: skipped
               "garbage"]
    c = s.stop()
    assert c["sm_mhz"] == 1957.5 and c["sm_max_mhz"] == 1965.0 and c["samples"] == 2 and c["reasons"] == ["sw_power_cap"]
    s2 = bench.ClockSampler(0)
    assert s2.stop()["reasons"] == ["nvidia-smi unavailable"]
    s.lines = ["0, 1300, 1965, 700.0, Active, Active, Not Active, Not Active"]
    assert s.stop()["reasons"] == ["hw_slowdown", "hw_thermal_slowdown"]


def test_committed_profile_evidence_is_parseable():
    """profiles/: the ncu launch list parses with tools/instep_summary.py's reader and names the BiCGSTAB kernels; the traffic
    file bench.py reads has an entry for every kernel of the roofline table (bench.ncu_traffic returns them only while
    csrc/cmfd_kernels.cu is the profiled source -- otherwise None, never a stale literal)."""
    import csv
    import json
    import os
    import bench
    root = bench.ROOT
    rows = [r for r in csv.reader(l for l in open(os.path.join(root, "profiles", "r02_launches.csv")) if l.startswith('"'))]
    names = {r[rows[0].index("Kernel Name")].split("(")[0] for r in rows[1:]}
    assert any("k_st" in n for n in names) and any("k_spmv_dot" in n for n in names) and any("k_update_xr" in n for n in names)
    rec = json.load(open(os.path.join(root, "profiles", "ncu_traffic.json")))
    for k in ("k_st", "k_spmv_dot", "k_spmv", "k_update_xr", "k_update_p", "k_residual", "k_fsrc_norms"):
        assert rec["kernels"][k]["dram_bytes_per_launch"] > 1e8
        v = bench.ncu_traffic(k)
        assert v is None or v == rec["kernels"][k]["dram_bytes_per_launch"]
