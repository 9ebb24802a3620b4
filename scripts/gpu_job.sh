( time python -m pytest tests/test_gpu_parity.py tests/test_gpu_shapes.py tests/test_gpu_mcmc.py -m gpu -q -x 2>&1 | grep -v "^$" ) > gpurun_out/r3l_pytest.log 2>&1
tail -6 gpurun_out/r3l_pytest.log | cut -c1-300
for c in c2 c3; do python scripts/kind_timing.py --config $c >> gpurun_out/r3l_kind.log 2>&1; done
python scripts/kind_timing.py --config c3 --chains 5 >> gpurun_out/r3l_kind.log 2>&1
python scripts/kind_timing.py --config c5 --chains 8 --pipeline flat+cache >> gpurun_out/r3l_kind.log 2>&1
grep -E "config|Error|error" gpurun_out/r3l_kind.log | cut -c1-420
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3l_launches_c3.csv python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r3l_ncu_c3.log 2>&1
python scripts/launch_summary.py gpurun_out/r3l_launches_c3.csv 2>&1 | head -5
