# two-lane overlap experiments: stream priority (pass 1 demoted) x pass-1 CTA length, without CUDA graphs
for prio in 0 1; do for k in 8 16 32; do
echo "GRAPH=0 PRIO=$prio K1_CTAS=$k"; CONESGPU_GRAPH=0 CONESGPU_PRIO=$prio CONESGPU_K1_CTAS=$k python tools/overlap_probe.py 512 60 2>&1 | grep "ms/step"
done; done
