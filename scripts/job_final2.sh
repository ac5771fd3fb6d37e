timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
bash scripts/job_x2.sh
