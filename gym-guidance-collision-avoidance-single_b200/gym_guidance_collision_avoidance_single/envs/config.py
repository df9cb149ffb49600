"""Parameters of the registered environments: same class, attribute names and values as the
reference's PKG/config.py:4-39, read at construction like PKG/SingleAircraftEnv.py:49-64.

Mutate the class attributes before constructing an environment (e.g. `Config.intruder_size = 80`),
exactly as with the reference.
"""
import math


class Config:
    # map and batch-independent sizes (PKG/config.py:6-12)
    window_width = 800
    window_height = 800
    intruder_size = 0          # the registered default; Simulators/config.py uses 80
    EPISODES = 1000
    G = 9.8
    tick = 30
    scale = 30

    # distances in pixels (PKG/config.py:15-19)
    minimum_separation = 555 / scale
    NMAC_dist = 150 / scale
    horizon_dist = 4000 / scale
    initial_min_dist = 3000 / scale
    goal_radius = 600 / scale

    # speeds in pixels per step (PKG/config.py:22-26)
    min_speed = 50 / scale
    max_speed = 80 / scale
    d_speed = 5 / scale
    speed_sigma = 2 / scale
    position_sigma = 10 / scale

    # heading, radians (PKG/config.py:29-30)
    d_heading = math.radians(5)
    heading_sigma = math.radians(2)

    # bank model parameters, unused by step (PKG/config.py:33-36)
    min_bank = -25
    max_bank = 25
    d_bank = 5
    bank_sigma = 4

    # StackEnv episode cap (PKG/config.py:39)
    max_steps = 1000
