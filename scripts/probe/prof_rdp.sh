set -e
CMD="python bench.py --workload rdp_scale --rdp-reads 524288 --no-cpu-baseline"
$CMD > gpurun_out/plain_rdp2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_mma_bound' -s 6 -c 2 -o gpurun_out/r2b_rdp_mma $CMD > gpurun_out/ncu4.log 2>&1
tail -2 gpurun_out/ncu4.log
