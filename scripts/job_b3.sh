python -m pytest tests -m gpu -q 2>&1 | tail -40
python bench.py --steps 10 --warmup 3 --cpu-sample 20000 > gpurun_out/bench_b3.json 2> gpurun_out/bench_b3.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_b3.json')); print(d['value'], d['ms_per_step']); print(json.dumps(d['e2e'])[:600]); print({k:round(v['ms_per_launch'],3) for k,v in d['roofline']['kernels'].items()})"
tail -5 gpurun_out/bench_b3.err
