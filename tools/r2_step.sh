# GPU box: parity of the restructured step, then timing
set -x
timeout 600 python -m pytest tests -m gpu -x -q 2>&1 | tail -15
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
GCA_BENCH_KERNEL_ONLY=1 timeout 300 python bench.py --steps 1000 --warmup 10 > gpurun_out/r2_step_kernel_only.json 2> gpurun_out/r2_step_kernel_only.err; tail -c 300 gpurun_out/r2_step_kernel_only.err
GCA_BENCH_KERNEL_ONLY=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_step_kernel_only_20.json 2>/dev/null
timeout 300 python tools/n0_bench.py > gpurun_out/r2_n0.json 2>&1
GCA_BENCH_KERNEL_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 100 --warmup 3 > gpurun_out/r2_ncu_l.log 2>&1
