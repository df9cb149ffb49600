export GCA_BENCH_KERNEL_ONLY=1
L=gym-guidance-collision-avoidance-single_b200/lib
for lib in libgca.so libgca_mb5.so libgca_mb6.so libgca_mb8.so; do for t in 8 16 32; do for st in 2 4; do
GCA_LIB=$PWD/$L/$lib GCA_TILE=$t GCA_STAGES=$st python bench.py --steps 1000 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('$lib tile',$t,'stages',$st, '%.3e'%d['value'], '%.3f'%d['roofline']['frac'], '%.1f us'%(1e3*d['ms_per_step']))"
done; done; done
