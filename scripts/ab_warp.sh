#!/bin/bash
# NOTE (round 2): the DSVC_TMA_CFG / DSVC_WARP_* knobs exist only in a tuning build:
#   DSVC_TUNE=1 python -c "import __graft_entry__ as g; g.build(force=True)"
# (the product library compiles one configuration and reads no environment variables).
# A/B of the staged warp kernel's tuning knobs at 1080p C=64 (run under gpurun).
out=gpurun_out/ab_warp.txt; : > $out
run() { echo "== $* $EXTRA" >> $out; env "$@" python scripts/prof_kernels.py --what feature --iters 40 $EXTRA 2>&1 | grep "warp_fwd" >> $out; }
EXTRA=""
run DSVC_TMA_CFG=15 DSVC_WARP_TAIL_SPLIT=2
run DSVC_TMA_CFG=19 DSVC_WARP_TAIL_SPLIT=2
run DSVC_TMA_CFG=20 DSVC_WARP_TAIL_SPLIT=2
run DSVC_TMA_CFG=15 DSVC_WARP_TAIL_SPLIT=2 DSVC_WARP_TAIL_PCT=70
run DSVC_TMA_CFG=15 DSVC_WARP_TAIL_SPLIT=2 DSVC_WARP_TAIL_PCT=150
run DSVC_TMA_CFG=15 DSVC_WARP_TAIL_SPLIT=1
run DSVC_TMA_CFG=13 DSVC_WARP_TAIL_SPLIT=2
cat $out
