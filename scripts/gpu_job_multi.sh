# usage: gpu_job_multi.sh TAG N [N ...]   (on a box with >= max N GPUs): bench lines at N GPUs incl. the
# multi-GPU workloads of BASELINE.json (c4, the c5 temperature sweep, the weak-scaled ladder)
tag=$1; shift
nvidia-smi -L | wc -l
for n in "$@"; do
  python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n > gpurun_out/${tag}_bench_c3_n$n.json 2> gpurun_out/${tag}_bench_n$n.err
  echo "rc $?"; tail -c 300 gpurun_out/${tag}_bench_n$n.err
  python - <<PY
import json
try:
    d=json.loads(open('gpurun_out/${tag}_bench_c3_n$n.json').read().strip().splitlines()[-1])
    print($n, {k:d.get(k) for k in ('value','ms_per_step','ladder_consistent','per_rank_ms_per_step')}, d['e2e']['value'])
    for k,v in d.get('other_workloads',{}).items(): print('  ',k,{kk:v.get(kk) for kk in ('value','ms_per_step','ladder_consistent','error')})
except Exception as e: print('no line', e)
PY
done
