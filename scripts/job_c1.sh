timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -40
bash scripts/variants.sh 1048576 2>&1 | grep -v "CK_WALKER" 
