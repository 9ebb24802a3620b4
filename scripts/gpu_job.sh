for t in 5 10 20; do
  for pdl in 0 1 0 1; do
    MOIRE_B200_PDL=$pdl python bench.py --workload c3 --temperatures $t --steps 100 --warmup 5 --no-cpu-baseline --no-other-workloads 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('c3 T=$t pdl=$pdl', round(d['ms_per_step'],4), 'ms/step', round(d['value'],1), 'steps/s')"
  done
done
