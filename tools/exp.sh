# scratch GPU run: the new variant's tests first, then the whole GPU suite, then its bench leg
set -x
python -m pytest tests -m gpu -x -q -k "mctsrnd or kernels_per_step or random_configurations" > gpurun_out/rnd_pytest.log 2>&1; tail -15 gpurun_out/rnd_pytest.log
python -m pytest tests -m gpu -q > gpurun_out/r1_pytest_gpu.log 2>&1; tail -8 gpurun_out/r1_pytest_gpu.log
python tools/mctsrnd_bench.py > gpurun_out/mctsrnd_bench.json 2> gpurun_out/mctsrnd_bench.err; cat gpurun_out/mctsrnd_bench.json; tail -3 gpurun_out/mctsrnd_bench.err
