#!/bin/bash
# A/B of programmatic dependent launch (MOIRE_B200_PDL=0/1): bench lines of c3 and c2 through the step graph
for w in c3 c2 c1; do
  for pdl in 0 1 0 1; do
    MOIRE_B200_PDL=$pdl python bench.py --workload $w --steps 100 --warmup 5 --no-cpu-baseline --no-other-workloads 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$w pdl=$pdl', round(d['ms_per_step'],4), 'ms/step', round(d['value'],1), 'steps/s e2e', round(d['e2e']['value'],1), d.get('ladder'))"
  done
done
