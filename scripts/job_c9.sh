timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
echo "== mixed bucketing"; timeout 300 python scripts/quick_mixed.py 65536 2>&1 | tail -8
echo "== mixed no bucketing"; CK_BUCKET=0 timeout 300 python scripts/quick_mixed.py 65536 2>&1 | tail -8
echo "== 1k"; timeout 300 python scripts/quick_bench.py 1048576 2>&1 | grep -E "ms/launch|pipelined"
