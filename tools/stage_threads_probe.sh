for nt in 0 1; do for t in 1 2 4; do echo "NT=$nt STAGE_THREADS=$t"; CONESGPU_STAGE_NT=$nt CONESGPU_STAGE_THREADS=$t timeout 300 python tools/latency_case.py 500 2>&1 | tail -1; done; done
