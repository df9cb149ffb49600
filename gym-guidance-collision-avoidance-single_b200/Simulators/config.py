"""Parameters of the MCTS-driven environment: names and values of the reference's
Simulators/config.py:4-55 (80 intruders, heading sigma 4 degrees, reward row scaled by 1/10)."""
import math


class Config:
    window_width = 800
    window_height = 800
    diagonal = 800
    intruder_size = 80
    EPISODES = 1000
    G = 9.8
    tick = 30
    scale = 30

    minimum_separation = 555 / scale
    NMAC_dist = 150 / scale
    horizon_dist = 4000 / scale
    initial_min_dist = 3000 / scale
    goal_radius = 600 / scale

    min_speed = 50 / scale
    max_speed = 80 / scale
    d_speed = 5 / scale
    speed_sigma = 2 / scale
    position_sigma = 10 / scale

    d_heading = math.radians(5)
    heading_sigma = math.radians(4)

    max_steps = 1000

    # reward row (Simulators/config.py:37-43)
    NMAC_penalty = -10 / 10
    conflict_penalty = -5 / 10
    wall_penalty = -5 / 10
    step_penalty = -0.01 / 10
    goal_reward = 10 / 10
    sparse_reward = False
    conflict_coeff = 0.00025

    # n nearest intruders in the observation of SingleAircraftDiscrete9HEREnv (Simulators/config.py:55)
    n = 4
