python -m pytest tests -m gpu -x -q 2>&1 | tail -3
export GCA_BENCH_KERNEL_ONLY=1
for f in 0; do
GCA_FUSE_FINISH=$f python bench.py --steps 1000 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('fuse',$f,'%.3e'%d['value'], '%.3f'%d['roofline']['frac'], '%.1f us'%(1e3*d['ms_per_step']))"
done
export GCA_LIB=$PWD/gym-guidance-collision-avoidance-single_b200/lib/libgca_timing.so
GCA_FUSE_FINISH=0 python tools/kstamps.py 2>&1 | tail -12
