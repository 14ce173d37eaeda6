timeout 1200 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; echo "bench rc=$?"
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
