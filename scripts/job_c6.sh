for rep in 1 2 3; do
echo "== default rep $rep"; timeout 300 python scripts/quick_bench.py 1048576 2>&1 | grep -E "^emit|pipelined"
echo "== emit1 rep $rep"; CK_LIB=$PWD/gpurun_variants/libck_emit1.so timeout 300 python scripts/quick_bench.py 1048576 2>&1 | grep -E "^emit|pipelined"
done
CK_LIB=$PWD/gpurun_variants/libck_emit1.so timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -3
