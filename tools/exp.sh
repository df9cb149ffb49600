python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export GCA_BENCH_KERNEL_ONLY=1
python bench.py --steps 1000 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.3e'%d['value'], '%.1f us'%(1e3*d['ms_per_step']), d['roofline']['kernels_ms'], '%.3f'%d['roofline']['frac'])"
