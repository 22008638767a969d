#!/bin/bash
# round-2 GPU session AA (1 GPU): lazy_adf option -- transparency test, parity suite, bench line (e2e with the deferred upload)
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python -m pytest tests/test_gpu_parity.py tests/test_results.py -m gpu -q -rfEs > $O/r2aa_pytest.log 2>&1; echo "pytest rc=$?"
tail -5 $O/r2aa_pytest.log | cut -c1-250
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2aa_bench_n1.json 2> $O/r2aa_bench_n1.err; echo "bench n1 rc=$?"
python -c "
import json; d=json.loads(open('$O/r2aa_bench_n1.json').read().strip().split('\n')[-1]); print('n1 ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], 'traffic', d['roofline']['traffic'], 'solve', d['solve']['seconds_to_keff_convergence'], d['solve']['outer_iterations'])"
tail -3 $O/r2aa_bench_n1.err
