timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -5
echo "== mixed"; CK_BUCKET=0 timeout 300 python scripts/quick_mixed.py 65536 2>&1 | tail -10
echo "== fanout"; timeout 300 python scripts/quick_fanout.py 4096 2>&1 | tail -8
