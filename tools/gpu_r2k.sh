#!/bin/bash
# round-2 GPU session K (1 GPU): value-leg anomaly (7 ms/step at N = 1 in session J), L2 hints on the grouped-load kernels, C4 full probe, full test suite
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 300 python tools/value_leg.py 2>&1 | tee $O/r2k_value_leg.txt
timeout 300 python tools/value_leg.py st_var=0 spmv_var=0 2>&1 | tee -a $O/r2k_value_leg.txt
timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > $O/r2k_bench_n1.json 2> $O/r2k_bench_n1.err; echo "bench n1 rc=$?"
python -c "
import json; d=json.loads(open('$O/r2k_bench_n1.json').read().strip().split('\n')[-1]); print('bench n1 ms/step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])"
L=tools/ab
timeout 900 python tools/ab_step.py h0=$L/lib_h0.so h1=$L/lib_h1.so h4=$L/lib_h4.so h5=$L/lib_h5.so --reps 2 2>&1 | tee $O/r2k_ab.txt
timeout 900 python tools/c4_full_gpu.py 2>&1 | tee $O/r2k_c4_full.txt
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=5 > $O/r2k_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $O/r2k_pytest.log
