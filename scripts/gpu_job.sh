set -x
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2q_pytest.log 2>&1
tail -12 gpurun_out/r2q_pytest.log | cut -c1-300
for c in c2 c3; do python scripts/kind_timing.py --config $c >> gpurun_out/r2q_kind.log 2>&1; done; grep config gpurun_out/r2q_kind.log
python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r2q_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_sf_cells' -s 3 -c 3 -o gpurun_out/r2q_c3_cells python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r2q_ncu_c3.log 2>&1
