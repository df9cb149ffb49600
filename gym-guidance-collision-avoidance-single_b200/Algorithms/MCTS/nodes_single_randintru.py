"""Drop-in for the reference's Algorithms/MCTS/nodes_single_randintru.py (the model Agent_RandInt.py plans with):
same class names and methods as nodes_single.py on the 6 N + 8 raw observation of
Simulators/SingleAircraftMCTSRandIntruderEnv - every intruder (x, y, vx, vy, speed, heading) is seen ((len - 8) // 6,
:47), may turn by up to 10 degrees with probability 0.1 after each sub-frame advance (:64-71), and the ownship speed
is clamped on itself (:74).  `move` and `rollout` run on the GPU (gca_mcts.cu, random_intruders model).

    state = SingleAircraftState(state=last_observation)
    root = SingleAircraftNode(state=state)
"""
from . import nodes_single as _base
from .config_single import Config  # noqa: F401


class SingleAircraftState(_base.SingleAircraftState):
    RANDOM_INTRUDERS = True


class SingleAircraftNode(_base.SingleAircraftNode):
    pass
