timeout 300 python scripts/diag_gate.py 2>&1 | tail -5
CK_LIB=$PWD/gpurun_variants/libck_hash1.so timeout 300 python scripts/diag_gate.py 2>&1 | tail -5
timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -15
echo "== old hash"; CK_LIB=$PWD/gpurun_variants/libck_hash1.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -8
