timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
CUDA_VISIBLE_DEVICES=0 timeout 600 python bench.py --steps 10 --warmup 3 --cpu-sample 20000 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo bench1_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_n1.json')); print(d['value'], d['ms_per_step']); print({k:v['ms_per_launch'] for k,v in d['roofline']['kernels'].items()}); print(d['e2e']['value'], d['e2e']['engine_level']['value'], d['e2e'].get('ceiling',{}).get('events_per_s'))"
bash scripts/job_x2.sh
