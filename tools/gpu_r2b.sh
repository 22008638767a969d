#!/bin/bash
# round-2 GPU session B (2 GPUs): the tests session A left red, library A/B on one GPU, multi-GPU step decomposition
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_shim_replay.py tests/test_gpu_fullsize.py tests/test_multi_gpu.py tests/test_z_all_decks.py -m gpu -q -rfEs -k "not lmw_refined_full" > $O/r2b_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $O/r2b_pytest.log
timeout 900 python tools/ab_step.py r01=tools/ab/lib_r01.so new=adpres_b200/libadpres_b200.so lb6=tools/ab/lib_lb6.so --reps 2 > $O/r2b_ab.txt 2>&1; cat $O/r2b_ab.txt
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
for k in weak strong; do
  timeout 300 $TR tools/mg_step.py $k 2>&1 | grep MG_STEP
  ADP_NO_FUSE_MAIL=1 timeout 300 $TR tools/mg_step.py $k 2>&1 | grep MG_STEP
done
ADP_NO_PEER=1 timeout 300 $TR tools/mg_step.py weak 2>&1 | grep MG_STEP
timeout 300 python tools/mg_step.py weak 2>&1 | grep MG_STEP
timeout 300 python tools/mg_step.py strong 2>&1 | grep MG_STEP
