# scratch GPU run
set -x
python -m pytest tests/test_gpu_mcts.py -m gpu -x -q > gpurun_out/mcts_pytest.log 2>&1; tail -5 gpurun_out/mcts_pytest.log
python tools/mctsrnd_bench.py > gpurun_out/mctsrnd_bench.json 2> gpurun_out/mctsrnd_bench.err; cat gpurun_out/mctsrnd_bench.json; tail -3 gpurun_out/mctsrnd_bench.err
