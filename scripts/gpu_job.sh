( time python -m pytest tests -m gpu -q -x 2>&1 | grep -v "^$" ) > gpurun_out/r3y_pytest.log 2>&1
tail -4 gpurun_out/r3y_pytest.log | cut -c1-300
for c in c2 c3 c5; do python scripts/kind_timing.py --config $c >> gpurun_out/r3y_kind.log 2>&1; done
python scripts/kind_timing.py --config c3 --chains 5 >> gpurun_out/r3y_kind.log 2>&1
python scripts/kind_timing.py --config c4 --chains 8 >> gpurun_out/r3y_kind.log 2>&1
grep -E "config|Error|error" gpurun_out/r3y_kind.log | cut -c1-250
