set -x
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2c_pytest.log 2>&1
tail -15 gpurun_out/r2c_pytest.log
