#!/bin/bash
# round-2 GPU session X (2 GPUs): the new 8-group z-slab test (quad nodal kernels on slabs), both data planes
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest "tests/test_multi_gpu.py::test_slabs_match_oracle[peer-LMW_refined]" "tests/test_multi_gpu.py::test_slabs_match_oracle[nccl-LMW_refined]" -m gpu -q -rfEs > $O/r2x_pytest.log 2>&1; echo "pytest rc=$?"
tail -30 $O/r2x_pytest.log | cut -c1-300
