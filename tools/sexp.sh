export GCA_BENCH_KERNEL_ONLY=1
L=$PWD/gym-guidance-collision-avoidance-single_b200/lib
for x in 0 1 2 3 4 0; do
  if [ $x = 0 ]; then unset GCA_LIB; else export GCA_LIB=$L/libgca_exp$x.so; fi
  timeout 300 python bench.py --steps 1000 --warmup 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('exp$x', d['ms_per_step'], d['roofline']['kernels_ms'])"
done
