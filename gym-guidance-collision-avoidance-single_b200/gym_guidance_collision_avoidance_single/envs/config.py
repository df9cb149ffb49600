"""Parameters of the registered environments (reference: PKG/config.py:4-39), read at construction like
PKG/SingleAircraftEnv.py:49-64.  `Config` is a plain class: mutate its attributes before constructing an environment
(e.g. `Config.intruder_size = 80`), exactly as with the reference.  It is assembled from tables (value, provenance)."""
import math

_SCALE = 30                      # reference units (feet, knots) per pixel


def _px(x):
    return x / _SCALE


_MAP = dict(window_width=800, window_height=800,          # PKG/config.py:6-7
            intruder_size=0,                               # the registered default (:9); Simulators/config.py uses 80
            EPISODES=1000, G=9.8, tick=30, scale=_SCALE)   # :10-13
_DISTANCES = dict(minimum_separation=_px(555), NMAC_dist=_px(150), horizon_dist=_px(4000),      # :15-19, pixels
                  initial_min_dist=_px(3000), goal_radius=_px(600))
_SPEEDS = dict(min_speed=_px(50), max_speed=_px(80), d_speed=_px(5), speed_sigma=_px(2),       # :22-26, pixels per step
               position_sigma=_px(10))
_HEADING = dict(d_heading=math.radians(5), heading_sigma=math.radians(2))                       # :29-30, radians
_BANK = dict(min_bank=-25, max_bank=25, d_bank=5, bank_sigma=4)                                 # :33-36, unused by step
_EPISODE = dict(max_steps=1000)                                                                 # :39, StackEnv episode cap

Config = type("Config", (object,), {**_MAP, **_DISTANCES, **_SPEEDS, **_HEADING, **_BANK, **_EPISODE})
