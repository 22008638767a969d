#!/bin/bash
# round-2 GPU session T (8 GPUs): the bench line at N = 8 (weak scaling, parity and strong legs inside) and at N = 4
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
for N in 8 4; do
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2951$N"
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 > $O/r2t_bench_n$N.json 2> $O/r2t_bench_n$N.err; echo "bench n$N rc=$?"
python -c "
import json; d=json.loads(open('$O/r2t_bench_n$N.json').read().strip().split('\n')[-1]); print('n$N ms/step', d['ms_per_step'], 'value', d['value'], 'e2e', d['e2e'] and d['e2e']['value'], 'parity', d.get('parity'), 'strong', d.get('strong') and (d['strong']['ms_per_step'], d['strong']['n1_ms_per_step'], d['strong']['speedup_vs_n1']))"
grep "^\[bench 0" $O/r2t_bench_n$N.err | tail -4
done
