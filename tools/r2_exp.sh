L=$PWD/gym-guidance-collision-avoidance-single_b200/lib
python tools/her_step_probe.py 2>/dev/null | tail -1
for x in 1 2 3; do GCA_LIB=$L/libgca_exp$x.so python tools/her_step_probe.py 2>/dev/null | tail -1; done
