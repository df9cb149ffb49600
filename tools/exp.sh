export GCA_BENCH_KERNEL_ONLY=1 GCA_GROUP=4 GCA_STAGES=2
for dbg in 8 0; do
GCA_DEBUG_SKIP=$dbg python bench.py --steps 600 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('skip',$dbg, '%.3e'%d['value'], '%.1f us'%(1e3*d['ms_per_step']))"
done
GCA_LIB=$PWD/gym-guidance-collision-avoidance-single_b200/lib/libgca_timing.so python tools/phase_timing.py 2>&1 | grep -v "finish by smid" | tail -16
