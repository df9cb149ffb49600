# GPU box: launch list of the bench command, then one full ncu capture of the streaming pass
export GCA_BENCH_KERNEL_ONLY=1
python bench.py --steps 100 --warmup 3 > gpurun_out/plain.log 2>&1 || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/launches_r1.csv python bench.py --steps 100 --warmup 3 > gpurun_out/ncu_l.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:step_intruders -s 30 -c 1 -f -o gpurun_out/prof_intruders python bench.py --steps 100 --warmup 3 > gpurun_out/ncu_f.log 2>&1
ncu --set full --clock-control none --cache-control none --import-source on -k regex:step_finish -s 30 -c 1 -f -o gpurun_out/prof_finish python bench.py --steps 100 --warmup 3 > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f.log gpurun_out/ncu_f2.log
