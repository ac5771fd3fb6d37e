echo "== default (16K)"; timeout 300 python scripts/quick_mixed.py 65536 2>&1 | grep -E "^walk|sum"
for t in 8192 4096 2048; do
echo "== $t"; CK_LIB=$PWD/gpurun_variants/libck_long$t.so timeout 300 python scripts/quick_mixed.py 65536 2>&1 | grep -E "^walk|sum"
done
