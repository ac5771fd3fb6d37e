python -m pytest tests/test_gpu_kafka.py -m gpu -q 2>&1 | tail -5
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu/exchange_parity.py 2>&1 | tail -15
