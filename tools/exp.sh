python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python - <<'PY'
import sys, json
sys.path[:0]=['.','gym-guidance-collision-avoidance-single_b200']
import bench
print(json.dumps(bench.bench_her(0)))
PY
