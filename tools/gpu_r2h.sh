#!/bin/bash
# round-2 GPU session H (1 GPU): C-kernel formulations A/B, C4 full-size probe
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/stm_ab.py 2 2>&1 | tee $O/r2h_stm_ab.txt
if [ -f tests/golden/c4_full_oracle_result.json ]; then timeout 900 python tools/c4_full_gpu.py 2>&1 | tee $O/r2h_c4_full.txt; fi
