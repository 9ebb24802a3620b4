#!/bin/bash
# on the GPU box: time every library under variants/ on the c3 workload (all 40 chains)
for f in variants/libmoire_*.so; do
  echo "== $f"
  MOIRE_B200_LIB=$PWD/$f python scripts/kind_timing.py --config ${CFG:-c3} --steps ${STEPS:-4} --warmup 3 2>&1 | grep '^{'
done
