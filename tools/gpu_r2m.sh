#!/bin/bash
# round-2 GPU session M (2 GPUs): mailbox wire format A/B at N = 2 (LL words + system fence in every CTA vs flag with acquire load)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
ADP_MG_PROFILE=0 timeout 300 $TR tools/mg_step.py weak 2>&1 | grep "MG_STEP" | tee $O/r2m_mg.txt
ADP_NO_MAIL_LL=1 timeout 300 $TR tools/mg_step.py weak 2>&1 | grep "MG_STEP\|PROFILE\|launches" | tee -a $O/r2m_mg.txt
ADP_NO_MAIL_LL=1 ADP_MG_PROFILE=0 timeout 300 $TR tools/mg_step.py strong 2>&1 | grep "MG_STEP" | tee -a $O/r2m_mg.txt
