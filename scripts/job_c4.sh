echo "== default (16K)"; timeout 300 python scripts/quick_mixed.py 65536 2>&1 | tail -12
for lib in gpurun_variants/libck_long65536.so; do
echo "== $lib"; CK_LIB=$PWD/$lib timeout 300 python scripts/quick_mixed.py 65536 2>&1 | tail -12
done
