timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 tests/multigpu/worker_sharded.py > gpurun_out/worker_sharded_2gpu.log 2>&1; grep -v "^W0\|OMP_NUM\|^\*\*\*" gpurun_out/worker_sharded_2gpu.log | tail -3
