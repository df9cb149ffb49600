set -x
GCA_HEAD_CTAS_PER_SM=4 timeout 300 python tools/kstamps.py > gpurun_out/r2_kstamps.log 2>&1
export GCA_BENCH_KERNEL_ONLY=1
for c in 3 4; do
GCA_HEAD_CTAS_PER_SM=$c timeout 300 python bench.py --steps 1000 --warmup 10 > gpurun_out/r2_h$c.json 2>/dev/null
done
python - <<'PY'
import json
for c in (3, 4):
    try:
        d = json.loads(open("gpurun_out/r2_h%d.json" % c).read().strip().splitlines()[-1])
        print(c, d["ms_per_step"], d["roofline"].get("kernels_ms"))
    except Exception as e:
        print(c, "failed", e)
PY
