python -m pytest tests -m gpu -q 2>&1 | tail -15
for w in fanout mixed reply; do
python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value']); print(d['workload_stats']); print({k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()}); print(d['cpu_baseline']['value'])"
tail -3 gpurun_out/bench_$w.err
done
