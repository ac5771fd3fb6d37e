timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tests/multigpu/exchange_parity.py 2>&1 | grep -v "^W0\|OMP_NUM\|^\*\*\*" | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29540 scripts/diag_exchange.py 2>&1 | grep -v "^W0\|OMP_NUM\|^\*\*\*" | tail -6
