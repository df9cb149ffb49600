# GPU box: parity, then timing of the step (kernel-only bench lines + launch list)
set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -5
timeout 300 python tools/kstamps.py > gpurun_out/r2_kstamps.log 2>&1; tail -9 gpurun_out/r2_kstamps.log
for c in 1 2 3 4; do
GCA_HEAD_CTAS_PER_SM=$c GCA_BENCH_KERNEL_ONLY=1 timeout 300 python bench.py --steps 1000 --warmup 10 > gpurun_out/r2_step_j$c.json 2>/dev/null
done
GCA_BENCH_KERNEL_ONLY=1 timeout 300 python bench.py --steps 20 --warmup 3 > gpurun_out/r2_step_20.json 2>/dev/null
GCA_BENCH_KERNEL_ONLY=1 timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 100 --warmup 3 > gpurun_out/r2_ncu_l.log 2>&1
python - <<'PY'
import json
for f in ("r2_step_j1", "r2_step_j2", "r2_step_j3", "r2_step_j4", "r2_step_20"):
    try:
        d = json.loads(open("gpurun_out/%s.json" % f).read().strip().splitlines()[-1])
        print(f, d["ms_per_step"], d["value"], d["roofline"].get("kernels_ms"), d["roofline"]["step"]["frac"])
    except Exception as e:
        print(f, "failed", e)
PY
