( time python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" ) > gpurun_out/r3k_pytest.log 2>&1
tail -6 gpurun_out/r3k_pytest.log | cut -c1-300
for c in c2 c3 c5; do python scripts/kind_timing.py --config $c >> gpurun_out/r3k_kind.log 2>&1; done
python scripts/kind_timing.py --config c3 --chains 5 >> gpurun_out/r3k_kind.log 2>&1
grep -E "config|Error|error" gpurun_out/r3k_kind.log | cut -c1-420
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3k_launches_c3.csv python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r3k_ncu_c3.log 2>&1
python scripts/launch_summary.py gpurun_out/r3k_launches_c3.csv 2>&1 | head -36
