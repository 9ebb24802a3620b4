( time python -m pytest tests -m gpu -q -x ) > gpurun_out/r3b_pytest.log 2>&1
tail -3 gpurun_out/r3b_pytest.log
for c in c2 c3 c5; do python scripts/kind_timing.py --config $c >> gpurun_out/r3b_kind.log 2>&1; done
grep -E "config" gpurun_out/r3b_kind.log | cut -c1-600
python scripts/class_census.py c2 40 > gpurun_out/r3b_census_c2.log 2>&1; cat gpurun_out/r3b_census_c2.log
python scripts/class_census.py c3 40 > gpurun_out/r3b_census_c3.log 2>&1; cat gpurun_out/r3b_census_c3.log
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r3b_launches_c2.csv python scripts/kind_timing.py --config c2 --steps 1 --warmup 1 > gpurun_out/r3b_ncu_c2.log 2>&1
python scripts/seq_summary.py gpurun_out/r3b_launches_c2.csv 30
