CMD="python bench.py --workload rdp_scale --rdp-reads 1048576 --no-cpu-baseline"
$CMD > gpurun_out/plain_rdp.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r2b_rdp_scale_launches.csv $CMD > gpurun_out/ncu3.log 2>&1
tail -1 gpurun_out/plain_rdp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['detail']['routing_last_slice'])"
