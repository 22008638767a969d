#!/bin/bash
# round-2 GPU session N (2 GPUs): one system fence per pushing CTA + acquire-load LL wait: multi-GPU tests, step A/B, profile
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/mg_step.py weak 2>&1 | grep "MG_STEP\|PROFILE\|launches" | tee $O/r2n_mg.txt
ADP_NO_MAIL_LL=1 ADP_MG_PROFILE=0 timeout 300 $TR tools/mg_step.py weak 2>&1 | grep "MG_STEP" | tee -a $O/r2n_mg.txt
ADP_MG_PROFILE=0 timeout 300 $TR tools/mg_step.py strong 2>&1 | grep "MG_STEP" | tee -a $O/r2n_mg.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rfEs --durations=3 > $O/r2n_pytest_mg.log 2>&1; echo "pytest multi-gpu rc=$?"
tail -6 $O/r2n_pytest_mg.log
