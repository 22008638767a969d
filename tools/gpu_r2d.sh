#!/bin/bash
# round-2 GPU session D (1 GPU): full suite, record-slimmed nodal kernels, bulk-copy k_st A/B + ncu, library A/B
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=8 > $O/r2d_pytest.log 2>&1; echo "pytest rc=$?"
tail -14 $O/r2d_pytest.log
timeout 600 python tools/nodal_fused_ab.py 2 2>&1 | tee $O/r2d_nodal_ab.txt
timeout 600 python tools/nodal_ab.py 4 8 2>&1 | tee -a $O/r2d_nodal_ab.txt
timeout 600 python tools/tma_ab.py 3 2>&1 | tee $O/r2d_tma_ab.txt
timeout 900 python tools/ab_step.py r01=tools/ab/lib_r01.so new=adpres_b200/libadpres_b200.so --reps 2 2>&1 | tee $O/r2d_ab.txt
cat > /tmp/tma_prof.py <<'PY'
import os, sys
sys.path.insert(0, os.getcwd())
from adpres_b200 import capi
import bench
p = bench.load_c2()
s = capi.Solver(p, **bench.CTL)
s.set_option("graphs", 0); s.set_option("bench_warmup", 2)
s.matrix_setup(1); s.init_flux(); s.outer_begin(capi.MODE_FORWARD)
s.outer_steps(capi.MODE_FORWARD, 1, 2)
for tma in (0, 1):
    s.set_option("st_tma", tma)
    print(tma, s.bench_kernel(1, 2))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_st' -o $O/r2d_kst python /tmp/tma_prof.py > $O/r2d_ncu_kst.log 2>&1; echo "ncu rc=$?"
