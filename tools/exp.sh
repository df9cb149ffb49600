# scratch GPU run
set -x
python -m pytest tests -m gpu -x -q -k "mctsrnd or random_configurations or kernels_per_step or random_intruder" > gpurun_out/rnd_pytest.log 2>&1; tail -5 gpurun_out/rnd_pytest.log
python tools/mctsrnd_bench.py > gpurun_out/mctsrnd_bench.json 2> gpurun_out/mctsrnd_bench.err; cat gpurun_out/mctsrnd_bench.json; tail -3 gpurun_out/mctsrnd_bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/mctsrnd_launches.csv python tools/mctsrnd_bench.py > gpurun_out/mctsrnd_ncu.log 2>&1
python tools/sanitize.py 2>&1 | grep mctsrnd
