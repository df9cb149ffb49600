"""Parameters of the MCTS forward model (reference: Algorithms/MCTS/config_single.py:4-64; the model's noise differs
from the env it plans for, Q25).  `Config` is a plain class whose attributes can be mutated before use, as in the
reference; here it is assembled from tables so that units and provenance sit next to the numbers."""
import math

_SCALE = 30                                   # reference units (feet / knots) per pixel


def _px(x):
    return x / _SCALE


_TABLE = [
    # map (config_single.py:6-13)
    ("window_width", 800), ("window_height", 800), ("diagonal", 800), ("intruder_size", 20),
    ("EPISODES", 1000), ("G", 9.8), ("tick", 30), ("scale", _SCALE),
    # distances, pixels (:16-20)
    ("minimum_separation", _px(555)), ("NMAC_dist", _px(150)), ("horizon_dist", _px(4000)),
    ("initial_min_dist", _px(3000)), ("goal_radius", _px(600)),
    # speeds, pixels per step; the model is noise-free in speed and position (:23-27)
    ("min_speed", _px(50)), ("max_speed", _px(80)), ("d_speed", _px(5)), ("speed_sigma", _px(0)),
    ("position_sigma", _px(0)),
    # heading, radians (:30-31)
    ("d_heading", math.radians(5)), ("heading_sigma", math.radians(2)),
    ("max_steps", 1000),
    # reward row, scaled by 1/10 (:37-43)
    ("NMAC_penalty", -10 / 10), ("conflict_penalty", -5 / 10), ("wall_penalty", -5 / 10),
    ("step_penalty", -0.01 / 10), ("goal_reward", 10 / 10), ("sparse_reward", False), ("conflict_coeff", 0.00025),
    ("n", 4),
    # search (:57-64): re-plan period, sub-frames per move, simulations, depth, (unused) exploration constant
    ("update_frame", 5), ("simulate_frame", 10), ("no_simulation", 100), ("search_depth", 3), ("C", 0.70710678118),
]

Config = type("Config", (object,), dict(_TABLE))
