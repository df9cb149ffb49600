# GPU box: parity, then pipeline-shape sweep of the step kernel (units per stage x ring depth x resident blocks per SM)
python -m pytest tests -m gpu -x -q 2>&1 | tail -5
export GCA_BENCH_KERNEL_ONLY=1
for bps in 7 14; do for g in 2 4; do for st in 2 4; do
GCA_BLOCKS_PER_SM=$bps GCA_GROUP=$g GCA_STAGES=$st python bench.py --steps 600 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bps',$bps,'group',$g,'stages',$st, '%.3e'%d['value'], '%.3f'%d['roofline']['frac'], '%.1f us'%(1e3*d['ms_per_step']))"
done; done; done
GCA_GROUP=4 GCA_STAGES=2 GCA_LIB=$PWD/gym-guidance-collision-avoidance-single_b200/lib/libgca_timing.so python tools/phase_timing.py 2>&1 | tail -14
