#!/bin/bash
# round-2 GPU session F (1 GPU): L2-reuse A/B (sweep reversal, evict-first hints), in-step DRAM traffic per kernel, quad nodal kernels
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python tools/ab_step.py base=tools/ab/lib_base.so rev=tools/ab/lib_rev.so hint=tools/ab/lib_hint.so revhint=tools/ab/lib_revhint.so --reps 2 2>&1 | tee $O/r2f_ab.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for v in base revhint; do
  timeout 600 ncu --cache-control none --clock-control none --metrics $M --profile-from-start off --csv --log-file $O/r2f_instep_$v.csv python tools/instep_prof.py tools/ab/lib_$v.so > $O/r2f_instep_$v.log 2>&1; echo "ncu $v rc=$?"
  python tools/instep_summary.py $O/r2f_instep_$v.csv | tee $O/r2f_instep_$v.txt
done
timeout 900 python tools/nodal_ab.py 8 4 2>&1 | tee $O/r2f_nodal_ab.txt
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=5 > $O/r2f_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $O/r2f_pytest.log
