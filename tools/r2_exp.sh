set -x
L=$PWD/gym-guidance-collision-avoidance-single_b200/lib
python tools/faithful_bench.py 2>/dev/null | tail -1 > gpurun_out/r2_faithful_base.json
for mb in 6 7; do GCA_LIB=$L/libgca_fm$mb.so python tools/faithful_bench.py 2>/dev/null | tail -1 > gpurun_out/r2_faithful_fm$mb.json; done
python tools/n0_bench.py 2>/dev/null | tail -1 > gpurun_out/r2_n0_base.json
for mb in 6 8; do GCA_LIB=$L/libgca_n0m$mb.so python tools/n0_bench.py 2>/dev/null | tail -1 > gpurun_out/r2_n0_m$mb.json; done
timeout 600 python -m pytest tests/test_gpu_relabel.py -x -q 2>&1 | tail -3
for f in r2_faithful_base r2_faithful_fm6 r2_faithful_fm7 r2_n0_base r2_n0_m6 r2_n0_m8; do echo $f; cut -c1-420 gpurun_out/$f.json; done
GCA_BENCH_KERNEL_ONLY=1 python bench.py --steps 1000 --warmup 10 | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('headline', d['ms_per_step'], d['roofline']['kernels_ms'])"
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
