N=${1:-8}
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 tests/multigpu/exchange_parity.py > gpurun_out/exchange_parity_${N}gpu.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*" gpurun_out/exchange_parity_${N}gpu.log | tail -3
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus $N --steps 8 --warmup 3 --cpu-sample 4000 > gpurun_out/r02_bench_${N}gpu.json 2> gpurun_out/r02_bench_${N}gpu.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_${N}gpu.json')); print(d['value'], d['ms_per_step'], d['workload_stats']['exchange_parity']); print(d['e2e']['value'], d['e2e']['engine_level']['value'], d['e2e']['ceiling']['events_per_s'], d['e2e']['frac_of_ceiling'], d['e2e']['all_publishes_seen_by_sinks'])"
grep -n "EngineError\|illegal\|Error\|Traceback" gpurun_out/r02_bench_${N}gpu.err | head -5
