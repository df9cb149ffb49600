set -x
export GCA_BENCH_KERNEL_ONLY=1
for cv in -1 40 60 72 85; do GCA_CARVEOUT=$cv timeout 300 python bench.py --steps 1000 --warmup 10 > gpurun_out/r2_sep_cv$cv.json 2>/dev/null; done
GCA_CARVEOUT=-1 timeout 300 python tools/kstamps.py > gpurun_out/r2_kstamps_own_def.log 2>&1
python - <<'PY'
import json
for cv in (-1, 40, 60, 72, 85):
    try:
        d = json.loads(open("gpurun_out/r2_sep_cv%d.json" % cv).read().strip().splitlines()[-1])
        print(cv, d["ms_per_step"], d["roofline"].get("kernels_ms"), d["gpu_launches"])
    except Exception as e:
        print(cv, "failed", e)
PY
