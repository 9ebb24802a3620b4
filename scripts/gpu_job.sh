python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r3t_plain_c3.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:'k_flat_eval' -s 20 -c 2 -o gpurun_out/r3t_c3_eval python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > gpurun_out/r3t_ncu_c3.log 2>&1
python scripts/ncu_digest.py gpurun_out/r3t_c3_eval.ncu-rep --first > gpurun_out/r3t_ncu_full_flat_eval_40chains_digest.txt 2>&1
head -60 gpurun_out/r3t_ncu_full_flat_eval_40chains_digest.txt
rm -f gpurun_out/r3t_c3_eval.ncu-rep
python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3t_bench_short.json 2> gpurun_out/r3t_bench_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r3t_ncu_launches_bench.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r3t_ncu_bench.log 2>&1
python scripts/launch_summary.py gpurun_out/r3t_ncu_launches_bench.csv > gpurun_out/r3t_launch_summary.txt 2>&1; head -40 gpurun_out/r3t_launch_summary.txt
gzip -f gpurun_out/r3t_ncu_launches_bench.csv
python bench.py > gpurun_out/r3t_bench_c3_n1.json 2> gpurun_out/r3t_bench.err; tail -c 600 gpurun_out/r3t_bench_c3_n1.json
