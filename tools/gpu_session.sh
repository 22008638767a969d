#!/bin/bash
# First GPU session of the next round in ONE gpurun call (the %XTAB kernels and the all-decks tests were written
# after round 1's GPU budget was spent and have only been checked on the host):
#   /usr/local/graft/bin/gpurun --timeout 2400 -- 'bash tools/gpu_session.sh'
# Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke rc=$?"
timeout 900 python -m pytest tests/test_xtab.py tests/test_z_all_decks.py -m gpu -q > gpurun_out/new_gpu_tests.log 2>&1
echo "new gpu tests rc=$?"; tail -5 gpurun_out/new_gpu_tests.log
timeout 300 python tools/xtab_time.py > gpurun_out/xtab_time.log 2>&1
echo "xtab_time rc=$?"; tail -3 gpurun_out/xtab_time.log
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/all_gpu_tests.log 2>&1
echo "all gpu tests rc=$?"; tail -3 gpurun_out/all_gpu_tests.log
timeout 600 python bench.py > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?"; cut -c1-300 gpurun_out/bench_n1.json
