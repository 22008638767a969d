#!/bin/bash
# round-2 GPU session U (1 GPU): the evidence for profiles/ -- launch list of a bench run, ncu --set full of every hot kernel,
# in-step DRAM traffic of the final library
cd "$(dirname "$0")/.."
O=gpurun_out
mkdir -p $O
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file $O/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e --no-solve > $O/r02_launches_bench.json 2> $O/r02_launches.err; echo "launch list rc=$?"
python tools/instep_summary.py $O/r02_launches.csv | tee $O/r02_launches_summary.txt
timeout 900 ncu --set full --clock-control none --profile-from-start off -o $O/r02_full python tools/ncu_one_each.py > $O/r02_full.log 2>&1; echo "ncu full rc=$?"
python tools/ncu_summary.py $O/r02_full.ncu-rep --traffic-json $O/ncu_traffic.json | tee $O/r02_ncu_full.txt
ls -la $O/r02_full.ncu-rep
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 600 ncu --cache-control none --clock-control none --metrics $M --profile-from-start off --csv --log-file $O/r02_instep_final.csv python tools/instep_prof.py > $O/r02_instep_final.log 2>&1; echo "ncu instep rc=$?"
python tools/instep_summary.py $O/r02_instep_final.csv | tee $O/r02_instep_final.txt
rm -f $O/r02_full.ncu-rep.tmp
# keep the pull below 64 MiB
if [ $(stat -c %s $O/r02_full.ncu-rep) -gt 40000000 ]; then rm $O/r02_full.ncu-rep; fi
