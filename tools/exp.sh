python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 1000 --warmup 10 > gpurun_out/bench_now.json 2> gpurun_out/bench_now.err; tail -c 3000 gpurun_out/bench_now.json; tail -3 gpurun_out/bench_now.err
