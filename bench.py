#!/usr/bin/env python
"""bench.py -- throughput of the ADPRES eigenvalue hot path on B200.

Metric (BASELINE.json): node-group unknowns per second per outer iteration, on the
configuration the metric is quoted on: IAEA-3D refined to 1 cm x 1 cm x 2 cm nodes
(170 x 170 x 190 mesh, 4 579 000 nodes, 2 groups = 9.158 M node-groups, SANM kernel),
iteration control `%ITER . 10 1e-5 1e-5 5 50` (nin = 10, nac = 5, nupd = 50; the stable choice, see CTL).

A "step" is one pass of the reference's `do p = 1, nout` loop body (mod_cmfd.f90:465-496):
G x (TSrc + bicg(nin)), FSrc, l2norm, [fiss_extrp], Integrate, k-eff, RelE, RelEg and, whenever
mod(p, nupd) == 0, the SANM nodal update + matrix_setup(0).  The K timed steps are the iterations
p = P0 .. P0+K-1 of a run that starts from flat flux at p = 1, with P0 placed so that the window holds a
nodal update (window(): K = 20, nupd = 50 -> p = 31..50); iterations 1 .. P0-1 run untimed.  Both arms
(and the e2e call) time the SAME iterations of the SAME mesh.

  value   device-resident: the K steps enqueued back to back (adp_outer_steps), CUDA events on the
          library's stream, max over ranks.
  e2e     the same K steps through the drop-in boundary the Fortran driver uses, with HOST
          (pinned) buffers: one outer() call = adp_set_xs (all cross sections H2D) +
          adp_set_state + adp_matrix_setup(1) + K x adp_outer_iter (scalars D2H, exit test on
          the host) + nodal updates + adp_get_state (flux, fission source D2H) + adp_powdis.
          h2d/d2h bytes are the totals of that call divided by K.
  roofline  every kernel class timed alone with CUDA events against its algorithmic bytes
          and the measured HBM peak; the dominant one (largest share of the step) is the headline.
  cpu_baseline  the C oracle (oracle/, a line-by-line port of the Fortran; no Fortran compiler
          exists in the image, so oracle/_ref cannot be built) on 1 host core -- the reference
          is serial -- on a bounded sample of the same full mesh: ms per outer iteration and ms per
          nodal update (the reference's two timed regions, ADPRES.f90:55-81).

N > 1 (torchrun): z-slab decomposition, one rank per GPU.  `value` is weak scaling: N copies of the core
stacked axially (190 planes per rank, same node sizes); halo planes are pushed by the kernels
over NVLink peer memory, scalar all-reduces go through peer-memory mailboxes (NCCL fallback).
Two more objects at N > 1: `parity` (BASELINE configs[2] sliced over the N ranks against the committed
CPU-oracle fixture) and `strong` (C2', 10.07 M nodes, on 1 GPU and sliced over the N GPUs).

`--impl reference` times the reference algorithm on the host CPU (oracle port, 1 thread) on the same
config: the full mesh, one outer() call, the same timed iterations.
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
    """Rank `rank`'s view of a problem made of `stack` copies of `base` on top of each other, every plane of `base`
    split into `zrefine` planes (same cross sections, zdel / zrefine).  With base = load_c2(sample_planes=19) (one plane
    per axial assembly): zrefine = 10 is the C2 core, stack = world the weak-scaling workload (one core copy per GPU);
    zrefine = 22, stack = 1 is C2' (170 x 170 x 418 = 10 073 800 nodes), the strong-scaling workload.
    The C ABI takes the GLOBAL sdata arrays, but a rank only ever reads its own z-slab (+2 ghost planes) of
    them, so the global arrays are allocated uncommitted (np.empty) and only that plane range is
    filled -- host memory per rank stays at the size of the slab instead of growing with N."""

    def __init__(self, base, world, rank, stack=None, zrefine=1):
        from adpres_b200.slab import slab_planes
        stack = world if stack is None else stack
        self.base, self.world = base, world
        for k in ("mode", "ng", "nmat", "nxx", "nyy", "npl", "bc", "xdel", "ydel", "ystag_smin", "ystag_smax",
                  "xstag_smin", "xstag_smax", "chi", "nout", "nin", "serc", "ferc", "nac", "kern"):
            setattr(self, k, getattr(base, k))
        core = base.nzz * zrefine                                  # planes of one core copy
        self.nzz = core * stack
        self.nnod = base.npl * self.nzz
        self.nupd = base.nupd
        self.zdel = np.tile(np.repeat(base.zdel / np.float64(zrefine), zrefine), stack)
        npl = base.npl
        self.ix = np.tile(base.ix[:npl], self.nzz)
        self.iy = np.tile(base.iy[:npl], self.nzz)
        self.iz = np.repeat(np.arange(1, self.nzz + 1, dtype=np.int32), npl)
        k0, k1 = slab_planes(self.nzz, world, rank)
        self.k0, self.k1 = k0, k1
        ka, kb = max(0, k0 - 2), min(self.nzz, k1 + 2)
        self.rows = slice(ka * npl, kb * npl)                     # global node range this rank touches
        planes = (np.arange(ka, kb) % core) // zrefine            # plane of the stack -> plane of `base`
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


def window(K, W, nupd):
    """First iteration number P0 of the timed steps p = P0 .. P0+K-1.  The reference's loop counts p from 1 and runs the
    nodal update when mod(p, nupd) == 0 (mod_cmfd.f90:490); with K < nupd a window right after the W warm-up steps would
    contain none, so it is placed to END on p = nupd (K = 20, nupd = 50: p = 31..50, exactly one nodal update + matrix_setup(0)
    inside, same for the GPU arm, the e2e call and the CPU arm).  Iterations 1 .. P0-1 run untimed (the last W are the warm-up)."""
    return W + 1 if K >= nupd - W else nupd - K + 1


def updates_in(P0, K, nupd):
    return sum(1 for q in range(P0, P0 + K) if q % nupd == 0)


def workload_config(world, K, W):
    """`config` of the JSON line -- the SAME dict for the b200 arm and for --impl reference."""
    nz = 190 * world
    P0 = window(K, W, CTL["nupd"])
    return {"workload": "IAEA-3D refined 1cm x 1cm x 2cm (BASELINE configs[1])" + (f", {world} cores stacked axially" if world > 1 else "") +
                        f": 170x170x{nz} mesh, {24100 * nz} nodes, 2 groups = {48200 * nz} node-groups, SANM kernel; 190 planes per GPU",
            "nin": CTL["nin"], "nac": CTL["nac"], "nupd": CTL["nupd"], "parallelism": f"z-slab x{world}",
            "timed_iterations": [P0, P0 + K - 1], "nodal_updates_in_timed_steps": updates_in(P0, K, CTL["nupd"]),
            "l2": "inputs larger than L2 (each kernel streams >= 290 MB per launch; 126 MB L2)"}


def run_oracle_window(P0, K):
    """The reference algorithm (oracle port, 1 thread -- the reference is serial) on the FULL C2 mesh: ONE outer() call
    of P0+K-1 iterations (serc = ferc = 0 never exits); the timed steps are its iterations P0 .. P0+K-1, read from the
    per-iteration time stamps the oracle keeps next to the reference's own two accumulators (CMFD / nodal update,
    ADPRES.f90:55-81).  Returns a dict."""
    from oracle import Oracle
    p = load_c2()
    nit = P0 + K - 1
    o = Oracle(p, nout=nit, nin=CTL["nin"], nac=CTL["nac"], nupd=CTL["nupd"], serc=0.0, ferc=0.0)
    t0 = time.perf_counter()
    o.outer(0)
    total = time.perf_counter() - t0
    wall, fdm, nod = o.trace_times()
    assert len(wall) == nit, (len(wall), nit)
    i0, i1 = P0 - 2, nit - 1                      # end of iteration P0-1 .. end of iteration P0+K-1
    w0 = wall[i0] if i0 >= 0 else wall[0] - (wall[1] - wall[0])
    f0 = fdm[i0] if i0 >= 0 else 0.0
    n0 = nod[i0] if i0 >= 0 else 0.0
    nupd_in = updates_in(P0, K, CTL["nupd"])
    dt = wall[i1] - w0
    units = p.nnod * p.ng * K
    return {"value": units / dt, "seconds": dt, "total_seconds": total, "ms_per_step": 1e3 * dt / K,
            "cmfd_ms_per_outer": 1e3 * (fdm[i1] - f0) / K,
            "nodal_ms_per_update": (1e3 * (nod[i1] - n0) / nupd_in) if nupd_in else None,
            "nodal_updates": nupd_in, "nnod": int(p.nnod), "ng": int(p.ng), "keff_after": o.state()["Ke"],
            "sample": f"full BASELINE configs[1] mesh ({p.nxx}x{p.nyy}x{p.nzz}, {p.nnod} nodes x {p.ng} groups), one outer() call of "
                      f"{nit} iterations from flat flux, timed iterations {P0}..{nit} ({nupd_in} nodal update(s) inside), "
                      f"nin={CTL['nin']} nac={CTL['nac']} nupd={CTL['nupd']}, {cpu_model()}"}


def run_oracle_sample(n_outer=3):
    """cpu_baseline of the b200 arm: a bounded sample (about 30 s) of the same workload on the same full mesh -- one
    warm-up + `n_outer` timed outer iterations and ONE nodal update (nodal_upd: SANM sweep + matrix_setup(0),
    mod_cmfd.f90:339-383), each timed on its own; combined with the step's mix (one update per nupd iterations)."""
    from oracle import Oracle
    p = load_c2()
    o = Oracle(p, nout=n_outer + 1, nin=CTL["nin"], nac=CTL["nac"], nupd=10 ** 9, serc=0.0, ferc=0.0)
    o.outer(0)
    wall, fdm, nod = o.trace_times()
    cmfd = (wall[-1] - wall[0]) / n_outer
    t0 = time.perf_counter()
    rc = o.nodal_upd(1)
    nodal = time.perf_counter() - t0
    per_step = cmfd + nodal / CTL["nupd"]
    return {"value": p.nnod * p.ng / per_step, "unit": UNIT, "cores": 1, "kind": "port",
            "cmfd_ms_per_outer": 1e3 * cmfd, "nodal_ms_per_update": 1e3 * nodal, "nodal_update_status": int(rc),
            "seconds": wall[-1] - wall[0] + nodal, "host_cores_available": host_cores(),
            "sample": f"full BASELINE configs[1] mesh ({p.nnod} nodes x {p.ng} groups): {n_outer} outer iterations after 1 warm-up "
                      f"+ 1 nodal update, value = unknowns / (cmfd + nodal/nupd), nin={CTL['nin']} nupd={CTL['nupd']}, "
                      f"C oracle (port of the Fortran; no Fortran compiler in the image), 1 thread, {cpu_model()}"}


def reference_arm(args, rank, world, emit):
    """--impl reference: the reference's CPU implementation of the path (oracle port of the Fortran; oracle/_ref cannot
    be built without a Fortran compiler), all the threads it can use = 1 (the reference is serial), on the b200 arm's
    config: same mesh, same %ITER, same timed iterations.  At N > 1 (weak scaling: N stacked cores) rank 0 times ONE of
    the N core copies -- the bounded sample of that workload; the cost per unknown of the serial code does not depend on N."""
    if rank != 0:
        return
    K, W = args.steps, args.warmup
    P0 = window(K, W, CTL["nupd"])
    r = run_oracle_window(P0, K)
    line = {
        "impl": "reference", "metric": METRIC, "value": r["value"], "unit": UNIT, "n_gpus": world, "steps": K,
        "warmup": W, "ms_per_step": r["ms_per_step"] * world, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic", "config": workload_config(world, K, W),
        "cpu_baseline": {"value": r["value"], "unit": UNIT, "cores": 1, "kind": "port", "sample": r["sample"],
                         "cmfd_ms_per_outer": r["cmfd_ms_per_outer"], "nodal_ms_per_update": r["nodal_ms_per_update"],
                         "seconds": r["seconds"], "total_seconds": r["total_seconds"], "host_cores_available": host_cores()},
        "e2e": {"value": r["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0, "keff_after": r["keff_after"],
    }
    emit(line)


def pinned(a):
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a.ravel(order="K"))).pin_memory()
    return t, t.numpy().reshape(a.shape, order="F" if a.flags.f_contiguous else "C")


def ncu_traffic(kernel):
    """dram read+write bytes per launch of `kernel` from the committed ncu --set full capture, ONLY if that capture was
    taken from the kernel source that is in the tree now (profiles/ncu_traffic.json records the md5 of csrc/cmfd_kernels.cu;
    tools/ncu_summary.py writes it).  Otherwise None: a literal would go stale with the next kernel edit."""
    import hashlib
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        rec = json.load(open(path))
        md5 = hashlib.md5(open(os.path.join(ROOT, "adpres_b200", "csrc", "cmfd_kernels.cu"), "rb").read()).hexdigest()
        if rec.get("cmfd_kernels_md5") != md5:
            return None
        return rec["kernels"].get(kernel, {}).get("dram_bytes_per_launch")
    except Exception:
        return None


def c3_parity(world, rank, local_rank, uid, capi):
    """Multi-GPU parity inside the bench line: BASELINE configs[2] (IAEA-3D, 4 x 4 nodes per assembly, 190 planes,
    183 160 nodes, nin = 4) sliced over the N ranks, run for the fixture's outer count and compared with
    the committed CPU-oracle result tests/golden/c3_oracle_result.json (the JSON is read; oracle/ is not imported)."""
    from adpres_b200.deck import Problem
    ref = json.load(open(os.path.join(ROOT, "tests", "golden", "c3_oracle_result.json")))
    with open(os.path.join(ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
        p = Problem.from_spec(json.load(fh)).refine(xdiv=ref["xdiv"], ydiv=ref["ydiv"], zdiv=ref["zdiv"])
    # nin of the fixture (4; the deck's default 2 is only marginally stable on this mesh, tools/order_probe.py)
    # exactly the oracle's outer count (its 1e-8 exit): the solution still moves by ~1e-5 in power from one nodal update to the
    # next (nupd = 104), so both sides must have seen the same number of updates -- a run that exits one update cycle later
    # (702 instead of 612 outers on two slabs) sits 1.3e-5 away in assembly power without being any less converged
    s = capi.Solver(p, device=local_rank, nranks=world, rank=rank, uid=uid, nout=ref["outers"], serc=0.0, ferc=0.0, nin=ref["nin"])
    t0 = time.perf_counter()
    rc, n = s.outer(0)
    rc = 0 if (rc == capi.STOP_MAXOUTER and n == ref["outers"]) else (rc or -1)
    dt = time.perf_counter() - t0
    ke = s.state()["Ke"]
    fasm, _, _ = s.asm_pow()                      # all-reduced over the slabs: every rank holds the whole map
    asm_ref = np.array(ref["asm_power"])
    nz = asm_ref > 0
    s.close()
    return {"config": "C3 (BASELINE configs[2]): 34x34x190, 183160 nodes, nin=%d nupd=104, z-slabs over %d GPUs, serc=ferc=%g" % (ref["nin"], world, ref["serc"]),
            "status": int(rc), "outers": int(n), "oracle_outers": int(ref["outers"]), "keff": ke, "keff_pcm": abs(ke - ref["keff"]) * 1e5,
            "asm_power_rel": float(np.abs(fasm[nz] / asm_ref[nz] - 1).max()), "seconds": dt,
            "reference": "tests/golden/c3_oracle_result.json (CPU oracle, tools/c3_oracle.py)"}


def main():
    # Only the JSON line may reach stdout: NCCL / torchrun banners go to stderr.
    real_stdout = os.dup(1)
    os.dup2(2, 1)

    def emit(line):
        os.write(real_stdout, (json.dumps(line) + "\n").encode())
    # a hang must become a traceback on stderr, not a silent driver timeout
    import faulthandler
    faulthandler.dump_traceback_later(int(os.environ.get("ADP_BENCH_WATCHDOG_S", "900")), exit=True)
    t_start = time.time()

    def log(msg):
        sys.stderr.write("[bench %s +%.1fs] %s\n" % (os.environ.get("RANK", "0"), time.time() - t_start, msg))
        sys.stderr.flush()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-solve", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="N > 1: skip the C3 parity solve and the C2' strong-scaling leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        # the CPU arm runs ~50 outer iterations of the full mesh on one core (3.5 min on the GPU boxes, 11 min in the build
        # container): its own, longer watchdog
        faulthandler.cancel_dump_traceback_later()
        faulthandler.dump_traceback_later(int(os.environ.get("ADP_BENCH_WATCHDOG_S", "3000")), exit=True)
        reference_arm(args, rank, world, emit)
        return

    import torch
    import torch.distributed as dist
    from adpres_b200 import capi
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- the hot path has no CPU fallback")
    torch.cuda.set_device(local_rank)

    def new_uid():
        """a fresh NCCL unique id for one more Solver spanning all ranks"""
        buf = (capi.C.c_ubyte * 128)()
        if rank == 0:
            assert capi.load().adp_comm_unique_id(buf) == 0
        t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
        dist.broadcast(t, 0)
        return bytes(t.cpu().tolist())
    uid = None
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        uid = new_uid()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # nvidia-smi takes driver locks while it starts up (kernel launches stall for tens of ms):
    # start the sampler now, long before the timed region, and let it reach its steady loop
    sampler = ClockSampler(local_rank)
    sampler.start()

    # ---------------------------------------------------------------- problem
    base19 = load_c2(sample_planes=19) if world > 1 else None
    p = load_c2() if world == 1 else SlabProblem(base19, world, rank, stack=world, zrefine=10)
    s = capi.Solver(p, device=local_rank, nranks=world, rank=rank, uid=uid, **CTL)
    units_per_step = p.nnod * p.ng
    N_own = (s.k1 - s.k0) * p.npl
    W, K = args.warmup, args.steps
    P0 = window(K, W, CTL["nupd"])
    n_upd = updates_in(P0, K, CTL["nupd"])

    def timed_steps(sv, units):
        """iterations 1 .. P0-1 untimed (>= W warm-up steps), then K timed steps p = P0 .. P0+K-1 on the device.
        Before that the whole sequence runs once untimed: the first nodal update of a context allocates its scratch
        (2.4 GB at this size) and loads its kernels, and on a box that has just been started that one-off cost was seen to
        reach 50 ms -- inside a 85 ms window.  The timed pass repeats exactly the same iterations from the same start."""
        sv.matrix_setup(1)
        sv.init_flux()
        sv.outer_begin(capi.MODE_FORWARD)
        sv.outer_steps(capi.MODE_FORWARD, 1, P0 + K - 1)
        sv.reset_nodal()
        sv.matrix_setup(1)
        sv.init_flux()
        sv.outer_begin(capi.MODE_FORWARD)
        sv.outer_steps(capi.MODE_FORWARD, 1, P0 - 1)
        barrier()
        l0 = sv.launch_count()
        sv.timer_start()
        rc_, ke_, ser_, fer_ = sv.outer_steps(capi.MODE_FORWARD, P0, K)
        ms_ = sv.timer_stop()
        barrier()
        assert rc_ == 0 and np.isfinite(ke_), (rc_, ke_)
        return max_over_ranks(ms_), sv.launch_count() - l0, ke_

    # ---------------------------------------------------------------- value (device resident)
    t_wait = time.time()
    while not sampler.lines and time.time() - t_wait < 5.0:
        time.sleep(0.05)
    n_before = len(sampler.lines)
    log("value leg")
    ms, launches, ke = timed_steps(s, units_per_step)
    log("value leg done: %.3f ms/step" % (ms / K))
    time.sleep(0.12)                       # let the 100 ms sampler take one more reading of the loaded state
    n_after = len(sampler.lines)
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
        if world > 1:
            # every rank reads back ITS slab (what d2h_bytes_per_step counts).  The library's default on several ranks also
            # gathers the other slabs into every rank's host arrays (option gather_results, for host code that consumes whole
            # sdata arrays); that all-gather is not part of the path measured here
            s.set_option("gather_results", 0)
        # the ADFs and sigf -- half of the cross-section bytes, read only by the nodal update and by PowDis -- go up behind the
        # outer iterations (option lazy_adf); the host arrays stay untouched until the call has returned its results
        s.set_option("lazy_adf", 1)
        for k in ("D", "sigr", "nuf", "sigf", "sigs", "dc", "exsrc"):
            hx[k] = host_buffer(getattr(p, k))
        hx["chi"] = np.asfortranarray(p.chi)
        # the state after iteration P0-1 (the warm-up of this leg), in host buffers
        s.matrix_setup(1)
        s.init_flux()
        s.outer_begin(capi.MODE_FORWARD)
        s.outer_steps(capi.MODE_FORWARD, 1, P0 - 1)
        if world == 1:
            st0 = s.state()
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

        log("e2e leg")
        one_call(max(W, 3), P0 - max(W, 3))       # warm-up of the host path (same state: dn is still zero before p = nupd)
        barrier()
        t0 = time.perf_counter()
        s.timer_start()
        ke_e2e = one_call(K, P0)
        ms_e2e_dev = s.timer_stop()
        barrier()
        wall = max_over_ranks(time.perf_counter() - t0)
        nd, G = N_own, p.ng
        h2d = 8 * (nd * (4 * G + G * G + 6 * G + G) + p.nmat * G) + 8 * nd * (G + 1)   # XS + dc + exsrc + chi ; f0, fs0
        d2h = 8 * nd * (G + 1) + 8 * nd + 8 * 18 * K                                     # f0, fs0 ; power ; scalars per step
        s.set_option("lazy_adf", 0)
        e2e = {"value": units_per_step * K / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d * world / K),
               "d2h_bytes_per_step": int(d2h * world / K), "ms_per_step": 1e3 * wall / K,
               "device_ms_per_step": ms_e2e_dev / K, "keff_after": ke_e2e, "nodal_updates_inside": n_upd,
               "what": "one outer() call through the C ABI with pinned host buffers: adp_set_xs + adp_set_state + "
                       "adp_matrix_setup + K x adp_outer_iter (+ adp_nodal_upd when mod(p, nupd) = 0) + adp_get_state + adp_powdis; "
                       "option lazy_adf: dc and sigf are uploaded on a second stream while the outer iterations run"
                       + ("; every rank uploads and reads back its own z-slab (gather_results = 0)" if world > 1 else "")}

    log("kernel timings")
    # ---------------------------------------------------------------- roofline of the dominant kernel
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (measured copy bandwidth)"
    else:
        peak, peak_src = 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"
    kern = {}
    rows = N_own
    G, nin = p.ng, CTL["nin"]
    # (name, bench id, algorithmic bytes/row per SURVEY.md 8(d) BiCGSTAB phase table, launches per outer iteration)
    table = (("k_st (C: s = r - alpha v on the fly, t = A s, (t,t), (t,s))", 1, 88.0, G * nin),
             ("k_spmv_dot (B: v = A p, (rs,v))", 0, 80.0, G * nin),
             ("k_spmv (plain v = A p, no dot product)", 8, 72.0, 0),
             ("k_update_xr (D: x, r update, rho)", 2, 56.0, G * nin),
             ("k_update_p (A: p update)", 3, 32.0, G * (nin - 1)),
             ("k_residual (P: source + residual)", 4, 8.0 * (7 + 1 + 1 + 2 * (G - 1) + 1 + 1) + 4.0, G),
             ("k_fsrc_norms (F: fission source + norms)", 5, 8.0 * (3 * G + 2), 1))
    for name, what, bytes_per_row, per_step in table:
        kms = s.bench_kernel(what, 20)
        kern[name] = {"ms": kms, "alg_bytes_per_row": bytes_per_row, "GBps": rows * bytes_per_row / (kms * 1e-3) / 1e9,
                      "frac": rows * bytes_per_row / (kms * 1e-3) / 1e9 / peak, "launches_per_step": per_step,
                      "ms_per_step": kms * per_step,
                      "ncu_dram_bytes_per_launch": ncu_traffic(name.split(" ")[0]) if (world == 1 and rows == 4579000) else None}
    nodal_ms = s.bench_kernel(7, 3)
    kern["nodal update (source + per-direction two-node kernels)"] = {
        "ms": nodal_ms, "alg_bytes_per_node": 8.0 * (41 * G + G ** 2), "launches_per_step": 1.0 / CTL["nupd"],
        "GBps": rows * 8.0 * (41 * G + G ** 2) / (nodal_ms * 1e-3) / 1e9,
        "frac": rows * 8.0 * (41 * G + G ** 2) / (nodal_ms * 1e-3) / 1e9 / peak, "ms_per_step": nodal_ms / CTL["nupd"]}
    # the dominant kernel = largest share of the step (agrees with the committed launch list under profiles/)
    dom_name = max((k for k in kern if "alg_bytes_per_row" in kern[k]), key=lambda k: kern[k]["ms_per_step"])
    dom = kern[dom_name]
    roofline = {"bound": "hbm", "kernel": dom_name.split(" ")[0], "achieved": dom["GBps"], "peak": peak, "unit": "GB/s",
                "frac": dom["frac"], "traffic": dom["ncu_dram_bytes_per_launch"],
                "traffic_unit": "bytes/launch (ncu dram read+write, profiles/ncu_traffic.json; null = no capture of the kernel source in the tree)",
                "alg_bytes_per_launch": rows * dom["alg_bytes_per_row"], "ms_per_launch": dom["ms"], "peak_source": peak_src,
                "spmv_on_72B_basis": {"k_spmv_dot": rows * SPMV_BYTES_PER_ROW / (kern["k_spmv_dot (B: v = A p, (rs,v))"]["ms"] * 1e-3) / 1e9 / peak,
                                      "k_spmv": kern["k_spmv (plain v = A p, no dot product)"]["frac"]},
                "kernels": kern}
    # whole outer iteration: SURVEY.md 8(d): bicg 8(8+32 nin) + TSrc 8(2(G-1)+4) + tail 8(3G+4)/G per node-group row,
    # plus the nodal updates inside the timed steps at 8(41G+G^2) per node
    row_bytes = 8.0 * (8 + 32 * CTL["nin"]) + 8.0 * (2 * (p.ng - 1) + 4) + 8.0 * (3 * p.ng + 4) / p.ng
    step_bytes = (units_per_step / world) * row_bytes + n_upd / K * (p.nnod / world) * 8.0 * (41 * G + G ** 2)
    step_gbps = step_bytes / (ms * 1e-3 / K) / 1e9
    roofline["outer_iteration"] = {"alg_bytes_per_row": row_bytes, "alg_bytes_per_step_per_gpu": step_bytes, "GBps_per_gpu": step_gbps,
                                   "frac": step_gbps / peak,
                                   "note": f"timed iterations {P0}..{P0 + K - 1} contain {n_upd} nodal update(s) + matrix_setup(0)"}

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
    clocks = sampler.stop()
    clocks["samples_during_value_region"] = max(0, n_after - n_before)
    s.close()
    del s

    # ---------------------------------------------------------------- N > 1: parity and strong scaling
    def extras():
        """C3 parity solve and C2' strong-scaling leg (N > 1); returns (parity, strong)"""
        log("C3 parity leg")
        parity = c3_parity(world, rank, local_rank, new_uid(), capi)
        log("strong-scaling leg")
        # C2' needs more inner iterations than C2: on its 0.91 cm planes nin = 10 is at the edge of the two-node iteration's
        # stability (tools/c2prime_probe.py), nin = 20 converges smoothly in every summation order
        CTLS = dict(CTL, nin=20)
        # strong scaling: C2' (170 x 170 x 418 = 10 073 800 nodes, BASELINE north star ">= 10 M nodes") on 1 GPU (rank 0
        # alone, the others wait) and sliced over the N ranks; same %ITER, same timed iterations
        n1_ms = None
        if rank == 0:
            s1 = capi.Solver(SlabProblem(base19, 1, 0, stack=1, zrefine=22), device=local_rank, **CTLS)
            s1.matrix_setup(1); s1.init_flux(); s1.outer_begin(capi.MODE_FORWARD)
            s1.outer_steps(capi.MODE_FORWARD, 1, P0 - 1)
            torch.cuda.synchronize()
            s1.timer_start()
            rc1, ke1, _, _ = s1.outer_steps(capi.MODE_FORWARD, P0, K)
            n1_ms = s1.timer_stop() / K
            assert rc1 == 0 and np.isfinite(ke1)
            s1.close()
        n1_ms = max_over_ranks(n1_ms if n1_ms is not None else 0.0)
        pN = SlabProblem(base19, world, rank, stack=1, zrefine=22)
        sN = capi.Solver(pN, device=local_rank, nranks=world, rank=rank, uid=new_uid(), **CTLS)
        log("strong-scaling leg: N ranks")
        msN, _, keN = timed_steps(sN, pN.nnod * pN.ng)
        log("strong-scaling leg done")
        sN.close()
        strong = {"config": "C2' IAEA-3D 1cm x 1cm x 0.91cm: 170x170x418 = 10073800 nodes x 2 groups, nin=20 nupd=50, "
                            f"z-slabs over {world} GPUs ({pN.k1 - pN.k0} planes on rank {rank})",
                  "ms_per_step": msN / K, "n1_ms_per_step": n1_ms, "speedup_vs_n1": n1_ms / (msN / K),
                  "unknowns_per_s": pN.nnod * pN.ng * K / (msN * 1e-3), "timed_iterations": [P0, P0 + K - 1], "keff_after": keN,
                  "limiter": "3 global reductions + 2 halo planes per BiCGSTAB sweep (mod_cmfd.f90:1229-1240): "
                             f"{(3 * CTLS['nin'] + 1) * 2 + 2} barrier points per step, fixed cost independent of the slab size"}


        return parity, strong
    # ---------------------------------------------------------------- CPU baseline (rank 0, N = 1)
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = run_oracle_sample()

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": K, "warmup": W,
        "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic", "config": workload_config(world, K, W),
        "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "cpu_baseline": cpu, "solve": solve,
        "keff_after_steps": ke,
    }
    if world > 1 and not args.no_extras:
        # the two extra legs must never cost the bench line: if one of them raises, or hangs (a rank that failed inside a
        # collective leaves the others waiting), the line goes out without it and says why
        done = threading.Event()

        def give_up():
            if not done.is_set():
                if rank == 0:
                    line["parity"] = line.get("parity") or {"error": "extra legs timed out"}
                    emit(line)
                os._exit(0)
        watchdog = threading.Timer(float(os.environ.get("ADP_BENCH_EXTRAS_S", "420")), give_up)
        watchdog.daemon = True
        watchdog.start()
        try:
            parity, strong = extras()
            line["parity"], line["strong"] = parity, strong
        except Exception as e:          # noqa: BLE001 -- reported in the line
            line["parity"] = {"error": "%s: %s" % (type(e).__name__, e)}
            log("extra legs failed: %r" % (e,))
            done.set()
            watchdog.cancel()
            if rank == 0:
                emit(line)
            os._exit(0)                 # the other ranks may sit in a collective of the failed leg
        done.set()
        watchdog.cancel()
    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
