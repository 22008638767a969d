#!/bin/bash
# round-2 GPU session W (2 GPUs): final validation -- smoke, the whole -m gpu suite, bench at N = 1 and N = 2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 1800 python -m pytest tests -m gpu -q -rfEs --durations=6 > $O/r2w_pytest.log 2>&1; echo "pytest rc=$?"
tail -12 $O/r2w_pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2w_bench_n1.json 2> $O/r2w_bench_n1.err; echo "bench n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2w_bench_n2.json 2> $O/r2w_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ("r2w_bench_n1","r2w_bench_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1])
        print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"] and "%.4g"%d["e2e"]["value"], "traffic", d["roofline"]["traffic"], "frac", d["roofline"]["frac"], "parity", d.get("parity") and (d["parity"]["keff_pcm"], d["parity"]["asm_power_rel"]), "strong", d.get("strong") and d["strong"]["speedup_vs_n1"])
    except Exception as e:
        print(f, "unreadable", e)
PY
