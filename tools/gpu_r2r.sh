#!/bin/bash
# round-2 GPU session R (2 GPUs): the whole -m gpu suite (multi-GPU cases, new C2' / C3 / C4 fixtures), bench at N = 1 and N = 2
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 600 python bench.py --steps 20 --warmup 5 > $O/r2r_bench_n1.json 2> $O/r2r_bench_n1.err; echo "bench n1 rc=$?"
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 20 --warmup 5 > $O/r2r_bench_n2.json 2> $O/r2r_bench_n2.err; echo "bench n2 rc=$?"
python - <<'PY'
import json
for f in ("r2r_bench_n1","r2r_bench_n2"):
    try:
        d=json.loads(open("gpurun_out/%s.json"%f).read().strip().split("\n")[-1])
        print(f, "ms/step %.3f"%d["ms_per_step"], "value %.4g"%d["value"], "e2e", d["e2e"] and "%.4g"%d["e2e"]["value"], "parity", d.get("parity"), "strong", d.get("strong") and {k:d["strong"][k] for k in ("ms_per_step","n1_ms_per_step","speedup_vs_n1")})
        print("   outer iteration frac", d["roofline"]["outer_iteration"]["frac"], "solve", d.get("solve") and (d["solve"]["seconds_to_keff_convergence"], d["solve"]["outer_iterations"], d["solve"]["keff"]), "cpu", d.get("cpu_baseline") and (d["cpu_baseline"]["cmfd_ms_per_outer"], d["cpu_baseline"]["nodal_ms_per_update"]))
        for k,v in d["roofline"]["kernels"].items(): print("   %-40s %.4f ms  frac %.3f"%(k[:40], v["ms"], v.get("frac",0)))
    except Exception as e:
        print(f, "unreadable", e)
PY
timeout 1800 python -m pytest tests -m gpu -q -rfEs --durations=8 > $O/r2r_pytest.log 2>&1; echo "pytest rc=$?"
tail -14 $O/r2r_pytest.log
