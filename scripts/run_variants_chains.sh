for f in variants/libmoire_*.so; do
  echo "== $f"
  for ch in 40 20 5; do
    MOIRE_B200_LIB=$PWD/$f python scripts/kind_timing.py --config c3 --chains $ch --steps 4 --warmup 3 2>&1 | grep '^{' | cut -c 1-120
  done
done
