#!/bin/bash
# round-2 GPU session L (2 GPUs): where the multi-rank step spends its time (per-launch-site profile at N = 1 and N = 2), C4 full probe
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/mg_step.py weak 2>&1 | grep -v "^W\|^\*\|OMP" | tee $O/r2l_mg_n1.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/mg_step.py weak 2>&1 | grep -v "^W\|^\*\|OMP" | tee $O/r2l_mg_n2.txt
timeout 300 $TR tools/mg_step.py strong 2>&1 | grep -v "^W\|^\*\|OMP" | tee $O/r2l_mg_n2_strong.txt
timeout 600 python tools/c4_full_gpu.py 2>&1 | tee $O/r2l_c4_full.txt
