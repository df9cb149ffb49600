# scratch GPU run
set -x
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "batch_split" 2>&1 | tail -4
