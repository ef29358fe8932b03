#!/bin/bash
# Device-resident step time of the 512-frame batch under the overlap knobs (two-lane default first).
# usage: bash tools/lane_sweep.sh > gpurun_out/lane_sweep.txt
B="python bench.py --steps 40 --no-sub-records --cpu-sample-frames 0 --latency-reps 5"
run() { echo "== $*"; env "$@" $B 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.4f  G pts/s %.1f' % (d['ms_per_step'], d['value']/1e9))"; }
run BENCH_WC=0
run BENCH_WC=0 CONESGPU_K1_CTAS=3
run BENCH_WC=0 CONESGPU_K1_CTAS=2
run BENCH_WC=0 CONESGPU_PRIO=1
run BENCH_WC=0 CONESGPU_PRIO=1 CONESGPU_K1_CTAS=3
run BENCH_WC=0 CONESGPU_STREAM_CTAS=8
run BENCH_WC=0 CONESGPU_FUSED_MASK=0
echo "== lanes 3"; BENCH_WC=0 $B --lanes 3 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.4f' % d['ms_per_step'])"
echo "== lanes 1"; BENCH_WC=0 $B --lanes 1 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ms/step %.4f' % d['ms_per_step'])"
