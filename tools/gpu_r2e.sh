#!/bin/bash
# round-2 GPU session E (2 GPUs): full -m gpu suite incl. multi-GPU cases, library A/B at N = 1, bench at N = 1 and N = 2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -q -rfEs --durations=10 > $O/r2e_pytest.log 2>&1; echo "pytest rc=$?"
tail -15 $O/r2e_pytest.log
timeout 900 python tools/ab_step.py r01=tools/ab/lib_r01.so new=adpres_b200/libadpres_b200.so --reps 2 2>&1 | tee $O/r2e_ab.txt
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2e_bench_n1.json 2> $O/r2e_bench_n1.err; echo "bench n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 600 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2e_bench_n2.json 2> $O/r2e_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ("r2e_bench_n1","r2e_bench_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1])
        print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"] and "%.4g"%d["e2e"]["value"], "parity", d.get("parity"), "strong", d.get("strong") and {k:d["strong"][k] for k in ("ms_per_step","n1_ms_per_step","speedup_vs_n1")})
        for k,v in d["roofline"]["kernels"].items(): print("   %-40s %.4f ms  frac %.3f"%(k[:40], v["ms"], v.get("frac",0)))
    except Exception as e:
        print(f, "unreadable", e)
PY
