timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -6
echo "== mixed, no bucketing"; CK_BUCKET=0 timeout 300 python scripts/quick_mixed.py 65536 2>&1 | grep -E "walk|sum"
echo "== mixed, bucketing"; timeout 300 python scripts/quick_mixed.py 65536 2>&1 | grep -E "walk|sum"
for lib in gpurun_variants/libck_minb6.so gpurun_variants/libck_minb8.so; do
echo "== mixed, bucketing, lib=$lib"; CK_LIB=$PWD/$lib timeout 300 python scripts/quick_mixed.py 65536 2>&1 | grep -E "walk|sum"
echo "== fanout lib=$lib"; CK_LIB=$PWD/$lib timeout 300 python scripts/quick_fanout.py 4096 2>&1 | grep -E "walk|fanout"
done
timeout 300 python bench.py --workload reply --steps 5 --warmup 3 > gpurun_out/bench_reply.json 2> gpurun_out/bench_reply.err; echo "reply rc=$?"
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_reply.json')); print('reply', d['value'], d['ms_per_step']); print(d['workload_stats']); print({k:(round(v['ms_per_launch'],3), v['launches']) for k,v in d['roofline']['kernels'].items()})"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:ck_plan_tool2_kernel -s 3 -c 1 -f -o gpurun_out/r02_plan2b python scripts/quick_bench.py 1048576 > gpurun_out/r02_ncu_plan2b.log 2>&1; tail -2 gpurun_out/r02_ncu_plan2b.log
