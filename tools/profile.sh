# GPU box: one full ncu capture of the MCTS playout kernel (bench must already exit 0 without ncu) + the raster kernel
cat > /tmp/mcts_only.py <<'PY'
import sys, json
sys.path[:0]=['.','gym-guidance-collision-avoidance-single_b200']
import bench
print(json.dumps(bench.bench_mcts(0, with_cpu=False)))
PY
python /tmp/mcts_only.py > gpurun_out/mcts_plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:mcts_playout -s 3 -c 1 -f -o gpurun_out/prof_mcts python /tmp/mcts_only.py > gpurun_out/ncu_mcts.log 2>&1
ncu --metrics smsp__sass_thread_inst_executed_op_dadd_pred_on.sum,smsp__sass_thread_inst_executed_op_dmul_pred_on.sum,smsp__sass_thread_inst_executed_op_dfma_pred_on.sum,gpu__time_duration.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active --clock-control none -k regex:mcts_playout -s 3 -c 1 --csv --log-file gpurun_out/mcts_ops.csv python /tmp/mcts_only.py > /dev/null 2>&1
cat gpurun_out/mcts_ops.csv | tail -6
tail -1 gpurun_out/mcts_plain.log
