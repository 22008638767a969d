#!/usr/bin/env python
"""bench.py -- throughput of the ADPRES eigenvalue hot path on B200.

Metric (BASELINE.json): node-group unknowns per second per outer iteration, on the
configuration the metric is quoted on: IAEA-3D refined to 1 cm x 1 cm x 2 cm nodes
(170 x 170 x 190 mesh, 4 579 000 nodes, 2 groups = 9.158 M node-groups, SANM kernel),
iteration control `%ITER . 10 1e-5 1e-5 5 50` (nin = 10, nac = 5, nupd = 50; the stable choice, see CTL).

A "step" is one pass of the reference's `do p = 1, nout` loop body (mod_cmfd.f90:465-496):
G x (TSrc + bicg(nin)), FSrc, l2norm, [fiss_extrp], Integrate, k-eff, RelE, RelEg and, whenever
mod(p, nupd) == 0, the SANM nodal update + matrix_setup(0).

  value   device-resident: K steps enqueued back to back (adp_outer_steps), CUDA events on the
          library's stream, max over ranks.
  e2e     the same K steps through the drop-in boundary the Fortran driver uses, with HOST
          (pinned) buffers: one outer() call = adp_set_xs (all cross sections H2D) +
          adp_set_state + adp_matrix_setup(1) + K x adp_outer_iter (scalars D2H, exit test on
          the host) + nodal updates + adp_get_state (flux, fission source D2H) + adp_powdis.
          h2d/d2h bytes are the totals of that call divided by K.
  roofline  the dominant kernel (BiCGSTAB SpMV + dot, k_spmv_dot) timed alone with CUDA events,
          algorithmic 72 B/row against the measured HBM peak.
  cpu_baseline  the C oracle (oracle/, a line-by-line port of the Fortran; no Fortran compiler
          exists in the image, so oracle/_ref cannot be built) on 1 host core -- the reference
          is serial -- on a bounded sample (same radial mesh, 19 of the 190 planes).

N > 1 (torchrun): z-slab decomposition, one rank per GPU, weak scaling: N copies of the core
stacked axially (190 planes per rank, same node sizes); halo planes are pushed by the kernels
over NVLink peer memory, scalar all-reduces go through peer-memory mailboxes (NCCL fallback).

`--impl reference` times the reference algorithm on the host CPU (oracle port, 1 thread).
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "node_group_unknowns_per_s_per_outer_iteration"
UNIT = "unknowns/s"
# %ITER: nin = 10 inner BiCGSTAB sweeps, extrapolation every 5, nodal update every 50 outers.  Chosen by
# measurement: on this 1 cm mesh the reference's default (nin = 2, nupd = 212) and nin = 2..5 with
# nupd = 50..100 make its two-node iteration unstable (ndmax > 1e3, the reference's own STOP, at the
# 1st-3rd nodal update); nin = 10 converges to k-eff 1.029069 in ~380 outer iterations.
CTL = dict(nin=10, nac=5, nupd=50, nout=1000000, serc=1e-5, ferc=1e-5)
SPMV_BYTES_PER_ROW = 72.0   # 7 coefficients + x + y, fp64


def load_c2(stack=1, sample_planes=None):
    """IAEA-3D (smpl/static/IAEA3Ds) with %GEOM lines 3/5/7 changed to `10 8*20`, `8*20 10`,
    `19*10` (BASELINE.json configs[1]: 1 cm x 1 cm x 2 cm nodes, 170 x 170 x 190).
    stack = n repeats the 19 axial assemblies n times (n cores on top of each other, same node
    sizes): the weak-scaling workload, one core copy per GPU.  (Refining the axial mesh n-fold
    instead is NOT a valid fixed-%ITER workload: the reference's two-node iteration goes unstable
    on 0.5 / 0.25 cm planes at nin = 10 -- ndmax > 1e3, its own STOP -- see DESIGN.md.)
    sample_planes = 19 coarsens the axial mesh to one plane per assembly (bounded CPU sample)."""
    import dataclasses
    from adpres_b200.deck import Problem
    with open(os.path.join(ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
        p = Problem.from_spec(json.load(fh))
    if stack > 1:
        p = dataclasses.replace(p, nz=p.nz * stack, zsize=np.tile(p.zsize, stack), zdiv=np.tile(p.zdiv, stack),
                                zpln=np.tile(p.zpln, stack))
    zdiv = [10] * p.nz if sample_planes is None else [sample_planes // 19] * p.nz
    return p.refine(xdiv=[10] + [20] * 8, ydiv=[20] * 8 + [10], zdiv=zdiv)


class SlabProblem:
    """Weak-scaling workload for rank `rank` of `world`: `world` copies of the C2 core stacked
    axially (190*world planes, same 1 cm x 1 cm x 2 cm nodes) == load_c2(stack=world).  The C ABI
    takes the GLOBAL sdata arrays, but a rank only ever reads its own z-slab (+2 ghost planes) of
    them, so the global arrays are allocated uncommitted (np.empty) and only that plane range is
    filled -- host memory per rank stays at the size of the slab instead of growing with N."""

    def __init__(self, base, world, rank):
        from adpres_b200.slab import slab_planes
        self.base, self.world = base, world
        for k in ("mode", "ng", "nmat", "nxx", "nyy", "npl", "bc", "xdel", "ydel", "ystag_smin", "ystag_smax",
                  "xstag_smin", "xstag_smax", "chi", "nout", "nin", "serc", "ferc", "nac", "kern"):
            setattr(self, k, getattr(base, k))
        self.nzz = base.nzz * world
        self.nnod = base.nnod * world
        self.nupd = base.nupd
        self.zdel = np.tile(base.zdel, world)
        npl = base.npl
        self.ix = np.tile(base.ix[:npl], self.nzz)
        self.iy = np.tile(base.iy[:npl], self.nzz)
        self.iz = np.repeat(np.arange(1, self.nzz + 1, dtype=np.int32), npl)
        k0, k1 = slab_planes(self.nzz, world, rank)
        self.k0, self.k1 = k0, k1
        ka, kb = max(0, k0 - 2), min(self.nzz, k1 + 2)
        self.rows = slice(ka * npl, kb * npl)                     # global node range this rank touches
        planes = np.arange(ka, kb) % base.nzz                     # plane of the stack -> plane of the single core
        src = (planes[:, None] * npl + np.arange(npl)[None, :]).reshape(-1)
        self.mat = np.empty(self.nnod, dtype=np.int32)
        self.mat[self.rows] = base.mat[src]
        for k in ("D", "sigr", "nuf", "sigf", "exsrc"):
            a = np.empty((self.nnod, base.ng), order="F")
            a[self.rows, :] = getattr(base, k)[src, :]
            setattr(self, k, a)
        self.sigs = np.empty((self.nnod, base.ng, base.ng), order="F")
        self.sigs[self.rows] = base.sigs[src]
        self.dc = np.empty((self.nnod, base.ng, 6), order="F")
        self.dc[self.rows] = base.dc[src]


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.lines = device, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1])); smax.append(float(f[2])); power.append(float(f[3]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(smax)), "power_w_max": float(max(power)),
                "samples": len(sm), "reasons": sorted(reasons)}


def host_cores():
    try:
        return len(os.sched_getaffinity(0))
    except Exception:
        return os.cpu_count() or 1


def cpu_model():
    try:
        for ln in open("/proc/cpuinfo"):
            if ln.startswith("model name"):
                return ln.split(":", 1)[1].strip()
    except Exception:
        pass
    return "unknown"


def run_oracle_sample(steps, warmup, sample_planes=19):
    """The reference algorithm (oracle port) on one host core: `warmup + steps` passes of the
    outer loop body on the bounded sample.  Returns (unknowns/s/iter, seconds, description)."""
    from oracle import Oracle
    p = load_c2(sample_planes=sample_planes)
    o = Oracle(p, nout=CTL["nout"], nin=CTL["nin"], nac=CTL["nac"], nupd=CTL["nupd"], serc=0.0, ferc=0.0)
    o.matrix_setup(1)
    o.init_flux()
    # drive the oracle's own outer() for warmup+steps iterations: serc = ferc = 0 never exits
    o.set_control(nout=warmup, nin=CTL["nin"], nac=CTL["nac"], nupd=CTL["nupd"], serc=0.0, ferc=0.0)
    if warmup > 0:
        o.outer(0)
    o.set_control(nout=steps, nin=CTL["nin"], nac=CTL["nac"], nupd=CTL["nupd"], serc=0.0, ferc=0.0)
    o.reset_times()
    t0 = time.perf_counter()
    o.outer(0)
    dt = time.perf_counter() - t0
    fdm, nod = o.times()
    units = p.nnod * p.ng * steps
    desc = (f"IAEA-3D 1 cm radial mesh, axial mesh coarsened to {p.nzz} planes ({p.nnod} nodes x {p.ng} groups), "
            f"{steps} outer iterations after {warmup} warm-up, nin={CTL['nin']} nac={CTL['nac']} nupd={CTL['nupd']}, "
            f"CMFD {fdm:.2f} s + nodal {nod:.2f} s, {cpu_model()}")
    return units / dt, dt, desc


def reference_arm(args, rank, emit):
    """--impl reference: the reference's CPU implementation of the path (oracle port of the
    Fortran; oracle/_ref cannot be built without a Fortran compiler), all the threads it can
    use = 1 (the reference is serial)."""
    if rank != 0:
        return
    val, dt, desc = run_oracle_sample(args.steps, args.warmup)
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "IAEA-3D refined 1cm x 1cm x 2cm (170x170x190, 4.579M nodes, 2 groups), SANM; "
                               "CPU arm runs a bounded sample of it", "nin": CTL["nin"], "nac": CTL["nac"], "nupd": CTL["nupd"]},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


def pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a.ravel(order="K"))).pin_memory()
    return t, t.numpy().reshape(a.shape, order="F" if a.flags.f_contiguous else "C")


def main():
    # Only the JSON line may reach stdout: NCCL / torchrun banners go to stderr.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, emit)
        return

    import torch
    import torch.distributed as dist
    from adpres_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        buf = (capi.C.c_ubyte * 128)()
        if rank == 0:
            assert capi.load().adp_comm_unique_id(buf) == 0
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().tolist())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # nvidia-smi takes driver locks while it starts up (kernel launches stall for tens of ms):
    # start the sampler now, long before the timed region, and let it reach its steady loop
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---------------------------------------------------------------- problem
    p = load_c2() if world == 1 else SlabProblem(load_c2(), world, rank)
    s = capi.Solver(p, device=local_rank, nranks=world, rank=rank, uid=uid, **CTL)
    units_per_step = p.nnod * p.ng
    N_own = (s.k1 - s.k0) * p.npl
    s.matrix_setup(1)
    s.init_flux()
    s.outer_begin(capi.MODE_FORWARD)
    W, K = args.warmup, args.steps

    # ---------------------------------------------------------------- value (device resident)
    s.outer_steps(capi.MODE_FORWARD, 1, W)
    t_wait = time.time()
    while not sampler.lines and time.time() - t_wait < 5.0:
        time.sleep(0.05)
    n_before = len(sampler.lines)
    barrier()
    l0 = s.launch_count()
    s.timer_start()
    rc, ke, ser, fer = s.outer_steps(capi.MODE_FORWARD, W + 1, K)
    ms = s.timer_stop()
    barrier()
    launches = s.launch_count() - l0
    time.sleep(0.12)                       # let the 100 ms sampler take one more reading of the loaded state
    n_after = len(sampler.lines)
    assert rc == 0 and np.isfinite(ke), (rc, ke)
    if world > 1:
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t.item())
    value = units_per_step * K / (ms * 1e-3)

    # ---------------------------------------------------------------- e2e (public API, host buffers)
    e2e = None
    if not args.no_e2e:
        keep, hx = [], {}
        cudart = torch.cuda.cudart()

        def host_buffer(a):
            """Page-locked host buffer with the contents of `a`: a pinned copy at N = 1; at N > 1 the
            (uncommitted) global array itself with only this rank's slab rows registered."""
            if world == 1:
                t, v = pinned(a)
                keep.append(t)
                return v
            a2 = a.reshape(a.shape[0], -1, order="F")
            for col in range(a2.shape[1]):
                seg = a2[p.rows, col]
                rc_ = cudart.cudaHostRegister(seg.ctypes.data, seg.nbytes, 0)
                assert int(rc_) == 0, f"cudaHostRegister failed: {rc_}"
            return a
        for k in ("D", "sigr", "nuf", "sigf", "sigs", "dc", "exsrc"):
            hx[k] = host_buffer(getattr(p, k))
        hx["chi"] = np.asfortranarray(p.chi)
        st0 = s.state() if world == 1 else None
        if world == 1:
            f0_h, fs0_h = host_buffer(st0["f0"]), host_buffer(st0["fs0"])
            ke0 = st0["Ke"]
        else:
            f0_h, fs0_h = np.empty((p.nnod, p.ng), order="F"), np.empty(p.nnod)
            ke_ = capi.C.c_double()
            s._chk(s.L.adp_get_state(s.h, capi._d(f0_h), capi._d(fs0_h), None, capi.C.byref(ke_)))
            ke0 = ke_.value
            host_buffer(f0_h); host_buffer(fs0_h)
        f0_out = host_buffer(np.empty((p.nnod, p.ng), order="F") if world > 1 else np.zeros((p.nnod, p.ng), order="F"))
        fs0_out = host_buffer(np.empty(p.nnod) if world > 1 else np.zeros(p.nnod))
        pw_out = host_buffer(np.empty(p.nnod) if world > 1 else np.zeros(p.nnod))
        L = s.L
        d = capi._d

        def one_call(nsteps, p_first):
            # what the patched Fortran outer() does: hand over sdata, iterate, take the results back
            s._chk(L.adp_set_xs(s.h, d(hx["D"]), d(hx["sigr"]), d(hx["nuf"]), d(hx["sigf"]), d(hx["sigs"]), d(hx["chi"]),
                                d(hx["dc"]), d(hx["exsrc"])))
            s._chk(L.adp_set_state(s.h, d(f0_h), d(fs0_h), capi.C.c_double(ke0)))
            s.matrix_setup(1)
            s.outer_begin(capi.MODE_FORWARD)
            for q in range(p_first, p_first + nsteps):
                s.outer_iter(capi.MODE_FORWARD, q)
                if q % CTL["nupd"] == 0:
                    s.nodal_upd(1)
            ke_ = capi.C.c_double()
            s._chk(L.adp_get_state(s.h, d(f0_out), d(fs0_out), None, capi.C.byref(ke_)))
            s._chk(L.adp_powdis(s.h, d(pw_out), 0))
            return ke_.value

        one_call(W, 1)
        barrier()
        t0 = time.perf_counter()
        s.timer_start()
        ke_e2e = one_call(K, W + 1)
        ms_e2e_dev = s.timer_stop()
        barrier()
        wall = time.perf_counter() - t0
        if world > 1:
            t = torch.tensor([wall], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            wall = float(t.item())
        nd, G = N_own, p.ng
        h2d = 8 * (nd * (4 * G + G * G + 6 * G + G) + p.nmat * G) + 8 * nd * (G + 1)   # XS + dc + exsrc + chi ; f0, fs0
        d2h = 8 * nd * (G + 1) + 8 * nd + 8 * 18 * K                                     # f0, fs0 ; power ; scalars per step
        e2e = {"value": units_per_step * K / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world / K),
               "d2h_bytes_per_step": int(d2h * world / K), "ms_per_step": 1e3 * wall / K,
               "device_ms_per_step": ms_e2e_dev / K, "keff_after": ke_e2e,
               "what": "one outer() call through the C ABI with pinned host buffers: adp_set_xs + adp_set_state + "
                       "adp_matrix_setup + K x adp_outer_iter (+ adp_nodal_upd every nupd) + adp_get_state + adp_powdis"}

    # ---------------------------------------------------------------- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    kern = {}
    rows = N_own
    G, nin = p.ng, CTL["nin"]
    # (name, bench id, algorithmic bytes/row per SURVEY.md 8(d) BiCGSTAB phase table, launches per outer iteration,
    #  dram bytes per launch from the committed ncu --set full capture profiles/r01_ncu_full.txt at the N = 1 size)
    table = (("k_st (C: s = r - alpha v on the fly, t = A s, (t,t), (t,s))", 1, 88.0, G * nin, 391.76e6),
             ("k_spmv_dot (B: v = A p, (rs,v))", 0, 80.0, G * nin, 360.72e6),
             ("k_spmv (plain v = A p, no dot product)", 8, 72.0, 0, 322.52e6),
             ("k_update_xr (D: x, r update, rho)", 2, 56.0, G * nin, 226.82e6),
             ("k_update_p (A: p update)", 3, 32.0, G * (nin - 1), 126.46e6),
             ("k_residual (P: source + residual)", 4, 8.0 * (7 + 1 + 1 + 2 * (G - 1) + 1 + 1) + 4.0, G, 491.38e6),
             ("k_fsrc_norms (F: fission source + norms)", 5, 8.0 * (3 * G + 2), 1, 286.86e6))
    for name, what, bytes_per_row, per_step, traffic in table:
        kms = s.bench_kernel(what, 20)
        kern[name] = {"ms": kms, "alg_bytes_per_row": bytes_per_row, "GBps": rows * bytes_per_row / (kms * 1e-3) / 1e9,
                      "frac": rows * bytes_per_row / (kms * 1e-3) / 1e9 / peak, "launches_per_step": per_step,
                      "ms_per_step": kms * per_step,
                      "ncu_dram_bytes_per_launch": traffic if (world == 1 and rows == 4579000) else None}
    nodal_ms = s.bench_kernel(7, 3)
    kern["nodal update (source + 3 node-direction + 3 surface launches)"] = {
        "ms": nodal_ms, "alg_bytes_per_node": 8.0 * (41 * G + G ** 2), "launches_per_step": 1.0 / CTL["nupd"],
        "GBps": rows * 8.0 * (41 * G + G ** 2) / (nodal_ms * 1e-3) / 1e9, "ms_per_step": nodal_ms / CTL["nupd"]}
    # the dominant kernel = largest share of the step (agrees with the committed launch list
    # profiles/r01_launches_summary.txt: k_st 20 %, k_spmv_dot 19 %)
    dom_name = max((k for k in kern if "alg_bytes_per_row" in kern[k]), key=lambda k: kern[k]["ms_per_step"])
    dom = kern[dom_name]
    roofline = {"bound": "hbm", "kernel": dom_name.split(" ")[0], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": dom["ncu_dram_bytes_per_launch"], "traffic_unit": "bytes/launch (ncu dram read+write)",
                "alg_bytes_per_launch": rows * dom["alg_bytes_per_row"], "ms_per_launch": dom["ms"], "peak_source": peak_src,
                "spmv_on_72B_basis": {"k_spmv_dot": rows * SPMV_BYTES_PER_ROW / (kern["k_spmv_dot (B: v = A p, (rs,v))"]["ms"] * 1e-3) / 1e9 / peak,
                                      "k_spmv": kern["k_spmv (plain v = A p, no dot product)"]["frac"]},
                "kernels": kern}
    # whole outer iteration: SURVEY.md 8(d): bicg 8(8+32 nin) + TSrc 8(2(G-1)+4) + tail 8(3G+4)/G per node-group row
    row_bytes = 8.0 * (8 + 32 * CTL["nin"]) + 8.0 * (2 * (p.ng - 1) + 4) + 8.0 * (3 * p.ng + 4) / p.ng
    step_gbps = (units_per_step / world) * row_bytes / (ms * 1e-3 / K) / 1e9
    roofline["outer_iteration"] = {"alg_bytes_per_row": row_bytes, "GBps_per_gpu": step_gbps, "frac": step_gbps / peak,
                                   "note": "includes the nodal updates that fall inside the timed steps"}

    # ---------------------------------------------------------------- seconds to k-eff convergence (N = 1)
    solve = None
    if world == 1 and not args.no_solve:
        s3 = capi.Solver(p, device=local_rank, **dict(CTL, nout=5000))
        s3.matrix_setup(1)                      # warm: allocations, module load
        t0 = time.perf_counter()
        rc3, n3 = s3.outer(0)                   # the reference's outer(): exit test every iteration, from f0 = 1
        torch.cuda.synchronize()
        dt3 = time.perf_counter() - t0
        solve = {"seconds_to_keff_convergence": dt3, "outer_iterations": n3, "keff": s3.state()["Ke"], "status": rc3,
                 "serc": CTL["serc"], "ferc": CTL["ferc"], "unknowns": units_per_step,
                 "what": "adp_outer(): full eigenvalue solve from flat flux incl. nodal updates, per-iteration exit test on the host"}
        s3.close()

    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        v, dt, desc = run_oracle_sample(steps=30, warmup=2)
        cpu = {"value": v, "unit": UNIT, "cores": 1, "kind": "port", "sample": desc, "seconds": dt,
               "host_cores_available": host_cores()}

    clocks = sampler.stop()
    clocks["samples_during_value_region"] = max(0, n_after - n_before)
    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
            "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": f"IAEA-3D refined 1cm x 1cm x 2cm" + (f", {world} cores stacked axially" if world > 1 else "") +
                                   f" ({p.nxx}x{p.nyy}x{p.nzz} mesh, {p.nnod} nodes, {p.ng} groups = {units_per_step} "
                                   f"node-groups), SANM kernel; {s.k1 - s.k0} planes per GPU",
                       "nin": CTL["nin"], "nac": CTL["nac"], "nupd": CTL["nupd"], "parallelism": f"z-slab x{world}",
                       "l2": "inputs larger than L2 (each kernel streams >= 290 MB per launch; 126 MB L2)",
                       "keff_after_steps": ke},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "solve": solve,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
