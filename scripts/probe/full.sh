python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "rc=$?"; tail -c 600 gpurun_out/bench_full.err
CMD="python bench.py --workload rdp_scale --rdp-reads 1048576 --no-cpu-baseline"
$CMD > gpurun_out/plain_rdp.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2b_rdp_scale_launches.csv $CMD > gpurun_out/ncu3.log 2>&1
echo done
