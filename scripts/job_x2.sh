python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu/exchange_parity.py > gpurun_out/x2.log 2>&1
grep -v "^W0\|^\[W\|warn" gpurun_out/x2.log | head -40
python -m pytest tests -m gpu -q 2>&1 | tail -15
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 --steps 10 --warmup 3 --cpu-sample 20000 > gpurun_out/bench_x2.json 2> gpurun_out/bench_x2.err; echo bench_rc=$?
python -c "
import json; d=json.load(open('gpurun_out/bench_x2.json')); print(d['value'], d['ms_per_step'], d['workload_stats']); print(json.dumps(d['e2e'])[:700])"
tail -5 gpurun_out/bench_x2.err
