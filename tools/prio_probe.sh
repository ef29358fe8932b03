for prio in 0 1; do for k in 4 8 16 32 64; do
echo "PRIO=$prio K1_CTAS=$k"; CONESGPU_PRIO=$prio CONESGPU_K1_CTAS=$k python tools/overlap_probe.py 512 60 2>&1 | grep "ms/step"
done; done
