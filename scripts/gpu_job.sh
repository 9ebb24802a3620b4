set -x
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r2a_pytest.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2a_smoke.log 2>&1
for c in c1 c2 c3 c5; do python scripts/kind_timing.py --config $c >> gpurun_out/r2a_kind.log 2>&1; done
python scripts/kind_timing.py --config c5 --chains 4 --steps 1 --warmup 1 > gpurun_out/r2a_plain_c5.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:k_flat_eval -s 34 -c 2 -o gpurun_out/r2a_c5_eval python scripts/kind_timing.py --config c5 --chains 4 --steps 1 --warmup 1 > gpurun_out/r2a_ncu_c5.log 2>&1
python scripts/kind_timing.py --config c2 --steps 1 --warmup 1 > gpurun_out/r2a_plain_c2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_flat_eval|k_pf_round' -s 80 -c 4 -o gpurun_out/r2a_c2_p python scripts/kind_timing.py --config c2 --steps 1 --warmup 1 > gpurun_out/r2a_ncu_c2.log 2>&1
tail -5 gpurun_out/r2a_pytest.log; cat gpurun_out/r2a_smoke.log | tail -2; cat gpurun_out/r2a_kind.log | grep config
