#!/usr/bin/env python
"""Two outer iterations of BASELINE configs[1] without CUDA graphs, for an IN-STEP per-kernel profile:
  ncu --cache-control none --clock-control none --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum \
      --profile-from-start off --csv --log-file out.csv python tools/instep_prof.py [lib.so]
The three metrics fit one pass, so nothing is replayed and the caches are not flushed between kernels: every launch
sees the L2 contents its predecessor left, as inside a real step.  tools/instep_summary.py sums the csv per kernel."""
import os, sys
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
from adpres_b200 import capi
if len(sys.argv) > 1:
    capi.LIB_PATH = os.path.abspath(sys.argv[1])
import bench
import torch
p = bench.load_c2()
s = capi.Solver(p, **bench.CTL)
s.set_option("graphs", 0)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 3)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
s.outer_steps(capi.MODE_FORWARD, 4, 2)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
