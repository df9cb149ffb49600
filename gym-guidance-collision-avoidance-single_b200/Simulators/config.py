"""Parameters of the environments under Simulators/ (reference: Simulators/config.py:4-55): 80 intruders, heading noise
of 4 degrees, the reward row scaled by 1/10, n = 4 nearest intruders in the Discrete{9,3}HER observation.  `Config` is a
plain mutable class like the reference's; it is assembled from tables (value, provenance) rather than written out."""
import math

_SCALE = 30


def _px(x):
    return x / _SCALE


_GEOMETRY = dict(window_width=800, window_height=800, diagonal=800,          # :6-8 (diagonal normalises distances)
                 intruder_size=80, EPISODES=1000, G=9.8, tick=30, scale=_SCALE)
_DISTANCES = dict(minimum_separation=_px(555), NMAC_dist=_px(150), horizon_dist=_px(4000),   # :15-19
                  initial_min_dist=_px(3000), goal_radius=_px(600))
_KINEMATICS = dict(min_speed=_px(50), max_speed=_px(80), d_speed=_px(5), speed_sigma=_px(2),  # :22-26
                   position_sigma=_px(10), d_heading=math.radians(5), heading_sigma=math.radians(4),   # :29-31
                   max_steps=1000)
_REWARDS = dict(NMAC_penalty=-10 / 10, conflict_penalty=-5 / 10, wall_penalty=-5 / 10,      # :37-43
                step_penalty=-0.01 / 10, goal_reward=10 / 10, sparse_reward=False, conflict_coeff=0.00025)
_OBSERVATION = dict(n=4)                                                                      # :55

Config = type("Config", (object,), {**_GEOMETRY, **_DISTANCES, **_KINEMATICS, **_REWARDS, **_OBSERVATION})
