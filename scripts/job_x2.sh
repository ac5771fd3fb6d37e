timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu/worker_sharded.py > gpurun_out/worker_sharded_2gpu.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*" gpurun_out/worker_sharded_2gpu.log | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu/exchange_parity.py > gpurun_out/exchange_parity_2gpu.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*" gpurun_out/exchange_parity_2gpu.log | tail -3
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-sample 20000 > gpurun_out/bench_x2.json 2> gpurun_out/bench_x2.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_x2.json')); print(d['value'], d['ms_per_step'], d['workload_stats']['exchange_parity']); print(d['e2e']['value'], d['e2e']['engine_level']['value'], d['e2e']['ceiling']['events_per_s'], d['e2e']['frac_of_ceiling'], d['e2e']['all_publishes_seen_by_sinks'])"
grep -n "EngineError\|illegal\|Error" gpurun_out/bench_x2.err | head -5
