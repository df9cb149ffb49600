# scratch GPU run
set -x
python -m pytest tests/test_gpu_mcts.py -m gpu -x -q > gpurun_out/mcts_pytest.log 2>&1; tail -15 gpurun_out/mcts_pytest.log
