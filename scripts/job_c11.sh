for rep in 1 2; do
echo "== default rep $rep"; timeout 300 python scripts/quick_bench.py 1048576 2>&1 | grep -E "^walk |^plan|^emit|pipelined"
echo "== minb8 rep $rep"; CK_LIB=$PWD/gpurun_variants/libck_minb8.so timeout 300 python scripts/quick_bench.py 1048576 2>&1 | grep -E "^walk |pipelined"
done
