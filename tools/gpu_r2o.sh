#!/bin/bash
# round-2 GPU session O (2 GPUs): push fence right after the boundary tiles; order sensitivity of the C3 card
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR tools/mg_step.py weak 2>&1 | grep "MG_STEP\|PROFILE\|launches" | tee $O/r2o_mg.txt
ADP_MG_PROFILE=0 timeout 300 $TR tools/mg_step.py strong 2>&1 | grep "MG_STEP" | tee -a $O/r2o_mg.txt
timeout 600 python tools/order_probe.py c3 2,3,4,6 0,600,900,1500 2>&1 | tee $O/r2o_c3_order.txt
timeout 900 python -m pytest tests/test_multi_gpu.py -m gpu -q -rfEs --durations=3 > $O/r2o_pytest_mg.log 2>&1; echo "pytest multi-gpu rc=$?"
tail -6 $O/r2o_pytest_mg.log
