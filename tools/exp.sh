# scratch GPU run: A/B of compile-time knobs on the headline step
set -x
python tools/ab_step.py 2>&1 | tail -3
GCA_LIB=$PWD/gym-guidance-collision-avoidance-single_b200/lib/ab/libgca_pdl.so python tools/ab_step.py 2>&1 | tail -3
GCA_LIB=$PWD/gym-guidance-collision-avoidance-single_b200/lib/ab/libgca_w2.so python tools/ab_step.py 2>&1 | tail -3
python tools/ab_step.py 2>&1 | tail -3
