timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -12
timeout 300 python scripts/quick_mixed.py 65536 2>&1 | tail -8
