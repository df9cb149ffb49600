# GPU box, end of round: full GPU test suite, smoke, bench (N=1), then (each only after its plain command exited 0)
# the ncu launch list of the kernel-only bench and one full capture of the dominant kernel.
set -x
python -m pytest tests -m gpu -q > gpurun_out/r1_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/r1_bench_reference.json 2> gpurun_out/r1_bench_reference.err
python bench.py > gpurun_out/r1_bench_n1.json 2> gpurun_out/r1_bench_n1.err || exit 1
GCA_BENCH_KERNEL_ONLY=1 python bench.py --steps 100 --warmup 3 > gpurun_out/r1_kernel_only.json 2> gpurun_out/r1_kernel_only.err || exit 1
GCA_BENCH_KERNEL_ONLY=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r1_launches.csv python bench.py --steps 100 --warmup 3 > gpurun_out/r1_ncu_l.log 2>&1
GCA_BENCH_KERNEL_ONLY=1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:step_intruders -s 30 -c 1 -f -o gpurun_out/r1_prof_intruders python bench.py --steps 100 --warmup 3 > gpurun_out/r1_ncu_f.log 2>&1
tail -c 400 gpurun_out/r1_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r1_mctsrnd_launches.csv python tools/mctsrnd_bench.py > gpurun_out/r1_mctsrnd_ncu.log 2>&1
python tools/mcts_experiment.py --envs 512 --episodes 1024 -s 90 --random-intruders > gpurun_out/r1_mcts_experiment_randint.txt 2>&1; tail -12 gpurun_out/r1_mcts_experiment_randint.txt
