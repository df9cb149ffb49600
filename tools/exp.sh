# scratch GPU run: every kernel at small / ragged sizes (compute-sanitizer is closed on this pool: plain run as a crash check)
set -x
timeout 600 python tools/sanitize.py > gpurun_out/san_plain.log 2>&1; echo rc=$?; tail -12 gpurun_out/san_plain.log
