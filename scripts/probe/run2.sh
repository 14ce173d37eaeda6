timeout 300 python -m pytest tests/test_stage_a_gpu.py -x -q -m gpu -k "plans_agree or random_models or edge or min_boot or ties or config3" 2>&1 | tail -3
python bench.py --steps 4 --warmup 3 --legs none --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['config']['routing_last_step'], d['roofline']['kernel_ms_per_launch'])"
PG_MMA_PROF=1 python bench.py --steps 1 --warmup 1 --legs none --no-cpu-baseline 2>&1 | grep k_mma_bound | tail -1
PG_MMA_PROF=1 python bench.py --workload rdp_scale --rdp-reads 2097152 --no-cpu-baseline 2> gpurun_out/rdp_prof.txt | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['detail'].get('routing'))"
grep k_mma_bound gpurun_out/rdp_prof.txt | tail -2 | head -1
