python -m pytest tests -m gpu -x -q 2>&1 | tail -25
python bench.py --steps 10 --warmup 3 --cpu-sample 20000 > gpurun_out/bench_b2.json 2> gpurun_out/bench_b2.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_b2.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'])[:1200]); print({k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"
tail -5 gpurun_out/bench_b2.err
ncu --set full --clock-control none --import-source on -k regex:ck_plan_tool2_kernel -s 3 -c 1 -f -o gpurun_out/r02_plan2 python scripts/quick_bench.py 1048576 > gpurun_out/r02_ncu_plan2.log 2>&1; tail -3 gpurun_out/r02_ncu_plan2.log
