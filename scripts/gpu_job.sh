set -x
python -m pytest tests/test_gpu_posterior.py -m gpu -q > gpurun_out/r2m_pytest_post.log 2>&1; tail -5 gpurun_out/r2m_pytest_post.log
python scripts/posterior_study.py c3s c5s > gpurun_out/r2m_posterior_study.json 2> gpurun_out/r2m_posterior_study.err; tail -3 gpurun_out/r2m_posterior_study.err
