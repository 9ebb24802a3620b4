# usage: gpu_job_multi.sh N [N ...]   (on a box with >= max N GPUs)
nvidia-smi -L | wc -l
if [ "$1" = "test" ]; then shift; ( time python -m pytest tests/test_gpu_multi.py -m gpu -q -x 2>&1 | grep -v "^$" ) > gpurun_out/r3u_pytest_multi.log 2>&1; tail -4 gpurun_out/r3u_pytest_multi.log | cut -c1-300; fi
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n > gpurun_out/r3u_bench_c3_n$n.json 2> gpurun_out/r3u_bench_n$n.err
  echo "rc $?"; tail -c 400 gpurun_out/r3u_bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/r3u_bench_c3_n$n.json').read().strip().splitlines()[-1])
    print($n, {k:d.get(k) for k in ('value','ms_per_step','ladder_consistent','per_rank_ms_per_step')}, d['e2e']['value'])
    for k,v in d.get('other_workloads',{}).items(): print('  ',k,{kk:v.get(kk) for kk in ('value','ms_per_step','chain_steps_per_sec','ladder_consistent','error')})
except Exception as e: print('no line', e)
PY
done
