#!/bin/bash
# One GPU-box call that produces everything profiles/ holds for a round (run under gpurun, 1 GPU):
#   gpurun --timeout 1500 -- 'bash scripts/profile_round.sh r01'
# Numbers printed by the runs under ncu are never bench values; the bench JSON comes from the plain run.
R=${1:-r01}
O=gpurun_out
mkdir -p $O
set -x
python -m pytest tests -m gpu -x -q > $O/${R}_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> $O/${R}_pytest_gpu.log
python bench.py > $O/${R}_bench_1gpu.json 2> $O/${R}_bench_1gpu.err; echo "bench rc=$?"
python bench.py --workload fanout > $O/${R}_bench_fanout.json 2> $O/${R}_bench_fanout.err; echo "fanout rc=$?"
python bench.py --workload reply > $O/${R}_bench_reply.json 2> $O/${R}_bench_reply.err; echo "reply rc=$?"
python bench.py --impl reference --steps 3 --warmup 1 > $O/${R}_bench_reference.json 2> $O/${R}_bench_reference.err; echo "reference rc=$?"
# launch list of the same bench command (per-launch times are cold-cache and serialised: compare shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/${R}_launches.csv \
    python bench.py --steps 2 --warmup 3 --events 262144 --cpu-sample 2000 > $O/${R}_launches_bench.log 2>&1
# full captures of the three heaviest kernels (one launch each, after warm-up launches)
for k in walk plan_tool emit; do
  ncu --set full --clock-control none --import-source on -k regex:ck_${k}_kernel -s $([ $k = walk ] && echo 4 || echo 3) -c 1 -f -o $O/${R}_${k} \
      python scripts/quick_bench.py 1048576 > $O/${R}_ncu_${k}.log 2>&1
done
ls -la $O
