set -x
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2t_pytest.log 2>&1
tail -12 gpurun_out/r2t_pytest.log | cut -c1-300
( time python -m pytest tests -m gpu -q ) > gpurun_out/r2t_pytest_again.log 2>&1
tail -4 gpurun_out/r2t_pytest_again.log | cut -c1-300
