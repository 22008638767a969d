#!/usr/bin/env python
"""BASELINE.json configs[2]: IAEA-3D radial map at 4 x 4 nodes per assembly, 190 planes
(34 x 34 x 190 = 183 160 nodes, 2 groups), STRONG scaling over z-slabs: the same problem on
1 / 2 / 4 / 8 GPUs.  With 366 k unknowns a kernel lasts 2-3 us, so the step is bound by launch and
all-reduce latency, not bandwidth (SURVEY 8(d) C3 says to report it as such).
usage: python tools/c3_strong.py            (1 GPU)
       torchrun --nproc-per-node N ... tools/c3_strong.py"""
import json, os, sys, time
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch
from adpres_b200 import capi
from adpres_b200.deck import Problem

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
uid = None
if world > 1:
    import torch.distributed as dist
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = (capi.C.c_ubyte * 128)()
    if rank == 0:
        assert capi.load().adp_comm_unique_id(buf) == 0
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    uid = bytes(t.cpu().tolist())
with open(os.path.join(ROOT, "tests", "golden", "IAEA3Ds.spec.json")) as fh:
    p = Problem.from_spec(json.load(fh)).refine(xdiv=[2] + [4] * 8, ydiv=[4] * 8 + [2], zdiv=[10] * 19)
ctl = dict(nin=2, nac=5, nupd=1000000, nout=100000)      # the reference's default inner work; no nodal update inside the timed steps
s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid, **ctl)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
W, K = 10, 400
s.outer_steps(capi.MODE_FORWARD, 1, W)
if world > 1:
    dist.barrier()
torch.cuda.synchronize()
s.timer_start()
rc, ke, ser, fer = s.outer_steps(capi.MODE_FORWARD, W + 1, K)
ms = s.timer_stop() / K
if world > 1:
    tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
    dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ms = float(tt.item())
if rank == 0:
    print("C3 strong scaling: %d nodes x %d groups on %d GPU(s): %.4f ms per outer iteration (nin=2) = %.3e unknowns/s/outer, Ke %.6f"
          % (p.nnod, p.ng, world, ms, p.nnod * p.ng / ms * 1e3, ke), flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
