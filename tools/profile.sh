export GCA_BENCH_KERNEL_ONLY=1
python bench.py --steps 100 --warmup 3 > gpurun_out/plain.log 2>&1 || exit 1
for k in step_finish step_own spawn_kernel; do
ncu --set full --clock-control none --cache-control none --import-source on -k regex:$k -s 30 -c 1 -f -o gpurun_out/prof_$k python bench.py --steps 100 --warmup 3 > gpurun_out/ncu_$k.log 2>&1
done
ls gpurun_out/*.ncu-rep
