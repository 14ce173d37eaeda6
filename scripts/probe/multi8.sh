python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 16 --warmup 3 --legs none --no-cpu-baseline > gpurun_out/scale8_weak.json 2> gpurun_out/scale8_weak.err; echo rc=$?
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --workload rdp_scale --no-cpu-baseline > gpurun_out/scale8_rdp.json 2> gpurun_out/scale8_rdp.err; echo rc=$?
tail -c 300 gpurun_out/scale8_weak.err; tail -c 300 gpurun_out/scale8_rdp.err
