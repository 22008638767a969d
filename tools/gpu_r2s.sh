#!/bin/bash
# round-2 GPU session S (2 GPUs): the tests touched since session R, bench N = 2 (parity at the oracle's outer count)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_fullsize.py tests/test_z_all_decks.py "tests/test_multi_gpu.py::test_slabs_match_oracle[peer-C3_fixture]" "tests/test_multi_gpu.py::test_slabs_match_oracle[nccl-C3_fixture]" tests/test_gpu_parity.py -m gpu -q -rfEs --durations=5 > $O/r2s_pytest.log 2>&1; echo "pytest rc=$?"
tail -8 $O/r2s_pytest.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2s_bench_n2.json 2> $O/r2s_bench_n2.err; echo "bench n2 rc=$?"
python -c "
import json; d=json.loads(open('$O/r2s_bench_n2.json').read().strip().split('\n')[-1]); print('n2 ms/step', d['ms_per_step'], 'parity', d.get('parity'), 'strong', d.get('strong') and d['strong']['speedup_vs_n1'])"
