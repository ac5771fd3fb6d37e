timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -8
timeout 300 python scripts/quick_mixed.py 65536 2>&1 | grep -E "walk|sum|status"
for w in fanout mixed; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value']); print(d['workload_stats']); print({k:(round(v['ms_per_launch'],3), v['launches']) for k,v in d['roofline']['kernels'].items()})"
tail -3 gpurun_out/bench_$w.err
done
