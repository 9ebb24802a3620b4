set -x
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2h_pytest.log 2>&1
tail -8 gpurun_out/r2h_pytest.log
for c in c2 c3 c5; do python scripts/kind_timing.py --config $c >> gpurun_out/r2h_kind.log 2>&1; done
python scripts/kind_timing.py --config c5 --chains 4 >> gpurun_out/r2h_kind.log 2>&1
python scripts/kind_timing.py --config c3 --chains 5 >> gpurun_out/r2h_kind.log 2>&1
grep config gpurun_out/r2h_kind.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2h_bench.json 2> gpurun_out/r2h_bench.err; cat gpurun_out/r2h_bench.json | cut -c1-400
python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r2h_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flat_eval -s 14 -c 1 -o gpurun_out/r2h_c3_eval python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r2h_ncu_c3.log 2>&1
