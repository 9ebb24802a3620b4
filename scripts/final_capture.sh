#!/bin/bash
# Round-end evidence on one B200: the bench line, the ncu launch list of the bench command, one full
# capture of the dominant kernel with source correlation, every named shape by kind.
# usage: scripts/final_capture.sh TAG      (files land in gpurun_out/TAG_*)
tag=${1:-r03}
out=gpurun_out
python bench.py > $out/${tag}_bench_c3_n1.json 2> $out/${tag}_bench.err; echo "bench rc $?"; tail -c 300 $out/${tag}_bench.err
python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-workloads > $out/${tag}_bench_short.json 2> $out/${tag}_bench_short.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $out/${tag}_launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-other-workloads > $out/${tag}_ncu_bench.log 2>&1
python scripts/launch_summary.py $out/${tag}_launches.csv > $out/${tag}_launch_summary.txt 2>&1; head -12 $out/${tag}_launch_summary.txt
gzip -f $out/${tag}_launches.csv
python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > $out/${tag}_plain_c3.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:k_flat_eval -s 16 -c 2 -f -o $out/${tag}_flat_eval \
    python scripts/kind_timing.py --config c3 --steps 1 --warmup 1 > $out/${tag}_ncu_c3.log 2>&1
python scripts/ncu_digest.py $out/${tag}_flat_eval.ncu-rep > $out/${tag}_ncu_full_flat_eval_40chains_digest.txt 2>&1; head -20 $out/${tag}_ncu_full_flat_eval_40chains_digest.txt
for c in c1 c2 c3 c4 c5; do
  extra=""; [ $c = c4 ] && extra="--chains 8"
  python scripts/kind_timing.py --config $c $extra 2>&1 | grep '"config"'
done > $out/${tag}_kind_timing_all_configs.log; cat $out/${tag}_kind_timing_all_configs.log | cut -c1-400
