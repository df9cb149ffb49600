GCA_BENCH_KERNEL_ONLY=1 timeout 300 python bench.py --steps 1000 --warmup 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', d['ms_per_step'], d['roofline']['kernels_ms'])"
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
