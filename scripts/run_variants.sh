#!/bin/bash
# on the GPU box: time every library under vbuild/ on a workload (default c3, all chains)
for f in vbuild/libmoire_*.so; do
  echo "== $f"
  MOIRE_B200_LIB=$PWD/$f python scripts/kind_timing.py --config ${CFG:-c3} --steps ${STEPS:-4} --warmup 3 ${EXTRA} 2>&1 | grep '^{'
done
