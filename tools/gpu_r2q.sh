#!/bin/bash
# round-2 GPU session Q (1 GPU): ncu of the G = 8 nodal update (quad kernels), isolated kernel times of the multi-rank forms
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none -k regex:'k_nodal' -o $O/r2q_nodal_g8 python tools/nodal_prof.py 8 2 > $O/r2q_ncu_nodal.log 2>&1; echo "ncu rc=$?"
python tools/ncu_summary.py $O/r2q_nodal_g8.ncu-rep | tee $O/r2q_nodal_g8.txt
timeout 600 python tools/ab_step.py new=adpres_b200/libadpres_b200.so --reps 1 2>&1 | tee $O/r2q_ab.txt
