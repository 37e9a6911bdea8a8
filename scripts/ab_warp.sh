#!/bin/bash
out=gpurun_out/ab_warp.txt; : > $out
run() { echo "== $* $EXTRA" >> $out; env "$@" python scripts/prof_kernels.py --what feature --iters 30 $EXTRA 2>&1 | grep "warp_fwd" >> $out; }
EXTRA=""
run DSVC_WARP_KNOBS=0
for k in 1 2 3 4 8 10; do run DSVC_WARP_KNOBS=$k; done
for p in 0 2 3; do run DSVC_TMA_PROMO=$p; done
cat $out
