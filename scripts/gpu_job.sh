for c in c2 c3 c5; do python scripts/kind_timing.py --config $c >> gpurun_out/r3q_kind.log 2>&1; done
python scripts/kind_timing.py --config c3 --chains 5 >> gpurun_out/r3q_kind.log 2>&1
python scripts/kind_timing.py --config c5 --chains 4 >> gpurun_out/r3q_kind.log 2>&1
python scripts/kind_timing.py --config c4 --chains 8 >> gpurun_out/r3q_kind.log 2>&1
grep -E "config|Error|error" gpurun_out/r3q_kind.log | cut -c1-250
( time python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" ) > gpurun_out/r3q_pytest.log 2>&1
tail -4 gpurun_out/r3q_pytest.log | cut -c1-300
python bench.py > gpurun_out/r3q_bench.json 2> gpurun_out/r3q_bench.err; python - <<'PY'
import json
d=json.loads(open('gpurun_out/r3q_bench.json').read().strip().splitlines()[-1])
print({k:d.get(k) for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['frac'], d['roofline']['ms_per_launch'], d['kernel_ms_per_step'], d['cpu_baseline']['value'])
PY
