timeout 600 python -m pytest tests -m gpu -q -x 2>&1 | tail -25
for w in fanout mixed reply; do
timeout 300 python bench.py --workload $w --steps 5 --warmup 3 > gpurun_out/bench_$w.json 2> gpurun_out/bench_$w.err; echo "$w rc=$?"
python -c "
import json,sys; d=json.load(open('gpurun_out/bench_$w.json')); print('$w', d['value'], d['ms_per_step'], d['e2e']['value']); print(d['workload_stats']); print({k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()}); print(d['cpu_baseline']['value'])"
tail -3 gpurun_out/bench_$w.err
done
timeout 300 python bench.py --steps 10 --warmup 3 --cpu-sample 20000 > gpurun_out/bench_b4.json 2> gpurun_out/bench_b4.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_b4.json')); print(d['value'], d['ms_per_step'], d['e2e']['value']); print({k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"
