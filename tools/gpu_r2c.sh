#!/bin/bash
# round-2 GPU session C (1 GPU): full suite, fused nodal kernels A/B + ncu, library A/B
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rfEs -k "not lmw_refined_full" > $O/r2c_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $O/r2c_pytest.log
timeout 600 python tools/nodal_fused_ab.py 2 4 2>&1 | tee $O/r2c_nodal_ab.txt
timeout 900 python tools/ab_step.py r01=tools/ab/lib_r01.so new=adpres_b200/libadpres_b200.so --reps 2 2>&1 | tee $O/r2c_ab.txt
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_nodal' -o $O/r2c_nodal python tools/nodal_prof2.py > $O/r2c_ncu_nodal.log 2>&1; echo "ncu rc=$?"
ls -la $O/*.ncu-rep
