#!/usr/bin/env python
"""Multi-GPU step time without the rest of bench.py: the weak-scaling workload (one C2 core copy per rank) or the strong
one (C2' sliced), K outer iterations with NO nodal update inside (p = 51..99), CUDA events, max over ranks.
torchrun --nproc-per-node N tools/mg_step.py [weak|strong] ;  env ADP_NO_FUSE_MAIL=1 / ADP_NO_PEER=1 select the older paths."""
import os, sys, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
import numpy as np
import torch, torch.distributed as dist
import bench
from adpres_b200 import capi
rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
uid = None
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    buf = (capi.C.c_ubyte * 128)()
    if rank == 0:
        assert capi.load().adp_comm_unique_id(buf) == 0
    t = torch.tensor(list(bytes(buf)), dtype=torch.uint8, device="cuda")
    dist.broadcast(t, 0)
    uid = bytes(t.cpu().tolist())
kind = sys.argv[1] if len(sys.argv) > 1 else "weak"
base19 = bench.load_c2(sample_planes=19)
p = bench.SlabProblem(base19, world, rank, stack=world, zrefine=10) if kind == "weak" else bench.SlabProblem(base19, world, rank, stack=1, zrefine=22)
s = capi.Solver(p, device=local, nranks=world, rank=rank, uid=uid, **bench.CTL)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 5)
res = []
for rep in range(3):
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s.timer_start()
    rc, ke, _, _ = s.outer_steps(capi.MODE_FORWARD, 51, 49)
    ms = s.timer_stop() / 49
    if world > 1:
        tt = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms = float(tt.item())
    res.append(ms)
# per-launch-site profile of a few steps without CUDA graphs (kernel + the gap / peer wait in front of it)
if os.environ.get("ADP_MG_PROFILE", "1") != "0":
    s.set_option("graphs", 0)
    s.outer_steps(capi.MODE_FORWARD, 51, 4)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    s.set_option("profile", 1)
    s.timer_start()
    s.outer_steps(capi.MODE_FORWARD, 51, 10)
    ms_ng = s.timer_stop() / 10
    rep = s.profile_report()
    s.set_option("profile", 0)
    s.set_option("graphs", 1)
    if rank == 0:
        print("PROFILE %s N=%d without graphs, with events: %.4f ms/step" % (kind, world, ms_ng))
        for name, cnt, ms in sorted(rep, key=lambda x: -x[2]):
            print("   %-28s %5d launches  %8.2f us each  %7.3f ms/step" % (name, cnt, 1e3 * ms / cnt, ms / 10), flush=True)
if rank == 0:
    print("MG_STEP %s N=%d fuse=%s peer=%s ms/step %s ke %.9f" % (kind, world, "0" if os.environ.get("ADP_NO_FUSE_MAIL") else "1",
          "0" if os.environ.get("ADP_NO_PEER") else "1", " ".join("%.4f" % x for x in res), ke), flush=True)
if world > 1:
    dist.barrier(); dist.destroy_process_group()
