# GPU box: per-warp phase timing, then one full ncu capture of the step kernel (bench must already exit 0 without ncu)
export GCA_GROUP=${GCA_GROUP:-4} GCA_STAGES=${GCA_STAGES:-2}
GCA_LIB=$PWD/gym-guidance-collision-avoidance-single_b200/lib/libgca_timing.so python tools/phase_timing.py > gpurun_out/phases.txt 2>&1
cat gpurun_out/phases.txt
export GCA_BENCH_KERNEL_ONLY=1
python bench.py --steps 20 --warmup 3 > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:step_kernel -s 5 -c 1 -f -o gpurun_out/prof_new python bench.py --steps 20 --warmup 3 > gpurun_out/ncu_f.log 2>&1
tail -3 gpurun_out/ncu_f.log
