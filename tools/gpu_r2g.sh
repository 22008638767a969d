#!/bin/bash
# round-2 GPU session G (1 GPU): L2 hint mask A/B, C2' iteration-control robustness, C4 full-size probe
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
L=tools/ab
timeout 1500 python tools/ab_step.py base=$L/lib_base.so rev=$L/lib_rev.so revA=$L/lib_revA.so revAB=$L/lib_revAB.so revACp=$L/lib_revACp.so revACc=$L/lib_revACc.so revADl=$L/lib_revADl.so revADs=$L/lib_revADs.so --reps 2 2>&1 | tee $O/r2g_ab.txt
timeout 900 python tools/c2prime_probe.py base=$L/lib_base.so rev=$L/lib_rev.so -- 10,50 10,100 10,200 10,304 15,50 20,50 2>&1 | tee $O/r2g_c2prime.txt
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
for v in rev revA; do
  timeout 600 ncu --cache-control none --clock-control none --metrics $M --profile-from-start off --csv --log-file $O/r2g_instep_$v.csv python tools/instep_prof.py $L/lib_$v.so > $O/r2g_instep_$v.log 2>&1; echo "ncu $v rc=$?"
  python tools/instep_summary.py $O/r2g_instep_$v.csv | tee $O/r2g_instep_$v.txt
done
if [ -f tests/golden/c4_full_oracle_result.json ]; then timeout 900 python tools/c4_full_gpu.py 2>&1 | tee $O/r2g_c4_full.txt; fi
