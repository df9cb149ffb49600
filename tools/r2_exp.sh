set -x
export GCA_BENCH_KERNEL_ONLY=1
for L in 0 32 64 128 256; do GCA_OWN_LEAD=$L timeout 300 python bench.py --steps 1000 --warmup 10 > gpurun_out/r2_lead$L.json 2>/dev/null; done
unset GCA_BENCH_KERNEL_ONLY
timeout 300 python tools/kstamps.py > gpurun_out/r2_kstamps_lead.log 2>&1
python - <<'PY'
import json
for L in (0, 32, 64, 128, 256):
    try:
        d = json.loads(open("gpurun_out/r2_lead%d.json" % L).read().strip().splitlines()[-1])
        print(L, d["ms_per_step"], d["roofline"].get("kernels_ms"))
    except Exception as e:
        print(L, "failed", e)
PY
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
