timeout 300 python -m pytest tests/test_stage_a_gpu.py -x -q -m gpu -k "plans_agree or random_models or tensor_core or config2 or config3" 2>&1 | tail -3
for P in 1 0; do
PG_PIPE=$P python bench.py --steps 8 --warmup 3 --legs none --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('pipe=$P', d['value'], d['e2e']['value'], d['config']['routing_last_step'], d['roofline']['kernel_ms_per_launch'])"
done
python bench.py --workload rdp_scale --rdp-reads 2097152 --no-cpu-baseline 2>/dev/null | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['e2e']['value'], d['detail']['routing_last_slice'])"
