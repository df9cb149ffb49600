"""Parameters of the MCTS forward model: names and values of the reference's
Algorithms/MCTS/config_single.py:4-64 (model noise differs from the env it plans for, Q25)."""
import math


class Config:
    window_width = 800
    window_height = 800
    diagonal = 800
    intruder_size = 20
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
    speed_sigma = 0 / scale
    position_sigma = 0 / scale

    d_heading = math.radians(5)
    heading_sigma = math.radians(2)

    max_steps = 1000

    NMAC_penalty = -10 / 10
    conflict_penalty = -5 / 10
    wall_penalty = -5 / 10
    step_penalty = -0.01 / 10
    goal_reward = 10 / 10
    sparse_reward = False
    conflict_coeff = 0.00025

    n = 4

    # MCTS algorithm (config_single.py:57-64)
    update_frame = 5
    simulate_frame = 10
    no_simulation = 100
    search_depth = 3
    C = 0.70710678118
