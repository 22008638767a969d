#!/bin/bash
# round-2 GPU session I (1 GPU): B-kernel formulations A/B (with the chosen C kernel), C4 full-size probe
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 python tools/spmv_ab.py 2 6 2>&1 | tee $O/r2i_spmv_ab.txt
if [ -f tests/golden/c4_full_oracle_result.json ]; then timeout 900 python tools/c4_full_gpu.py 2>&1 | tee $O/r2i_c4_full.txt; fi
