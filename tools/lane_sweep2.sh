#!/bin/bash
# Static SM partition between pass 1 and the per-frame kernel of the other lane.
B="python bench.py --steps 40 --no-sub-records --cpu-sample-frames 0 --latency-reps 5"
run() { echo "== $*"; env "$@" $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.4f  G pts/s %.1f' % (d['ms_per_step'], d['value']/1e9))"; }
run BENCH_WC=0
run BENCH_WC=0 CONESGPU_FRAME_CTAS=2
run BENCH_WC=0 CONESGPU_FRAME_CTAS=1
run BENCH_WC=0 CONESGPU_FRAME_CTAS=2 CONESGPU_K1_CTAS=3
run BENCH_WC=0 CONESGPU_FRAME_CTAS=2 CONESGPU_K1_CTAS=2
run BENCH_WC=0 CONESGPU_FRAME_CTAS=1 CONESGPU_K1_CTAS=3
run BENCH_WC=0 CONESGPU_FRAME_CTAS=3 CONESGPU_K1_CTAS=3
