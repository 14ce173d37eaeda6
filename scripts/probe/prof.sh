set -e
CMD="python bench.py --steps 2 --warmup 3 --legs none --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2b_launches_bench_steps2_warmup3.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_guess_bm|k_mma_meta|k_mma_bound|k_light|k_resolve' -s 30 -c 6 -o gpurun_out/r2b_final_full $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
