# GPU box, round 2: bench lines (both arms), then - each only after its plain command exited 0 - the ncu launch list of
# the kernel-only bench, full captures of the step's two kernels, and the f64 operation counts of the MCTS kernels.
set -x
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py --impl reference --steps 100 --warmup 3 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err || { tail -5 gpurun_out/r2_bench_n1.err; exit 1; }
python bench.py --steps 20 --warmup 3 > gpurun_out/r2_bench_n1_steps20.json 2> /dev/null
GCA_BENCH_KERNEL_ONLY=1 python bench.py --steps 100 --warmup 3 > gpurun_out/r2_kernel_only.json 2> gpurun_out/r2_kernel_only.err || exit 1
GCA_BENCH_KERNEL_ONLY=1 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 100 --warmup 3 > gpurun_out/r2_ncu_l.log 2>&1
GCA_BENCH_KERNEL_ONLY=1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:step_intruders -s 30 -c 1 -f -o gpurun_out/r2_prof_intruders python bench.py --steps 100 --warmup 3 > gpurun_out/r2_ncu_f.log 2>&1
GCA_BENCH_KERNEL_ONLY=1 ncu --set full --clock-control none --cache-control none --import-source on -k regex:step_finish -s 30 -c 1 -f -o gpurun_out/r2_prof_finish python bench.py --steps 100 --warmup 3 > gpurun_out/r2_ncu_f2.log 2>&1
cat > /tmp/mcts_only.py <<'PY'
import sys, json
sys.path[:0]=['.','gym-guidance-collision-avoidance-single_b200']
import bench
print(json.dumps(bench.bench_mcts(0, with_cpu=False)))
print(json.dumps(bench.bench_mctsrnd(0)))
PY
python /tmp/mcts_only.py > gpurun_out/r2_mcts_plain.log 2>&1 || exit 1
M=smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active
ncu --metrics $M --clock-control none -k regex:mcts_search_kernel -c 1 --csv --log-file gpurun_out/r2_mcts_search_ops.csv python /tmp/mcts_only.py > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:mcts_playout_rnd_lane -s 2 -c 1 --csv --log-file gpurun_out/r2_mctsrnd_model_ops.csv python /tmp/mcts_only.py > /dev/null 2>&1
ncu --metrics $M --clock-control none -k regex:mcts_playout_packed -s 3 -c 1 --csv --log-file gpurun_out/r2_mcts_ops.csv python /tmp/mcts_only.py > /dev/null 2>&1
tail -4 gpurun_out/r2_mcts_search_ops.csv gpurun_out/r2_mctsrnd_model_ops.csv gpurun_out/r2_mcts_ops.csv | cut -c1-400
tail -c 600 gpurun_out/r2_bench_n1.err
