python -m pytest tests -m gpu -x -q 2>&1 | tail -5
export GCA_BENCH_KERNEL_ONLY=1
for t in ${TILES:-8 16 32}; do for st in ${STAGES:-2 4}; do
GCA_TILE=$t GCA_STAGES=$st python bench.py --steps 1000 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tile',$t,'stages',$st, '%.3e'%d['value'], '%.3f'%d['roofline']['frac'], '%.1f us'%(1e3*d['ms_per_step']))"
done; done
