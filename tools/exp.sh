# scratch GPU run: parity suite and smoke on the current tree
set -x
python -m pytest tests -m gpu -q > gpurun_out/r1_pytest_gpu.log 2>&1; tail -3 gpurun_out/r1_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python tools/mctsrnd_bench.py 2>/dev/null | cut -c1-220
