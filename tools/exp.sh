# scratch GPU run: one full capture of the turn / observation pass of the random-intruder env
set -x
ncu --set full --clock-control none --cache-control none --import-source on -k regex:turn_obs -s 3 -c 1 -f -o gpurun_out/r1_prof_turn python tools/mctsrnd_bench.py > gpurun_out/r1_ncu_turn.log 2>&1; tail -3 gpurun_out/r1_ncu_turn.log
