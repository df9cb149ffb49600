"""ctypes mirror of the structs of include/gca.h that the CPU oracle takes, and the bench workload's configuration.

TEST INFRASTRUCTURE, NOT PRODUCT.  The oracle keeps its OWN definitions so that nothing under oracle/ imports the
product package (bench.py's reference arm runs with the oracle alone: no libgca.so, no gca_b200).  The product's
definitions live in gca_b200/abi.py; tests/test_host_logic.py checks that both mirror the header identically.
"""
import ctypes as C


class GcaConfig(C.Structure):                       # gca_config
    _fields_ = [(n, C.c_double) for n in (
        "window_width", "window_height",
        "minimum_separation", "nmac_dist", "initial_min_dist", "goal_radius",
        "min_speed", "max_speed", "d_speed", "speed_sigma",
        "d_heading", "heading_sigma",
        "ob_window_width", "ob_window_height", "ob_min_speed", "ob_max_speed",
        "r_nmac", "r_conflict", "r_wall", "r_goal", "r_default")] + [(n, C.c_int32) for n in (
            "shaped_default", "action_kind", "obs_kind", "wall_kind", "max_steps", "time_limit", "random_start",
            "nearest_n")] + [("ob_diagonal", C.c_double), ("conflict_coeff", C.c_double), ("goal_margin", C.c_double),
                             ("shaped_nearest", C.c_int32), ("intruder_turns", C.c_int32),
                             ("position_drift", C.c_double), ("turn_prob", C.c_double), ("turn_max_deg", C.c_double)]


class GcaHostState(C.Structure):                    # gca_host_state
    _fields_ = [(n, C.c_void_p) for n in ("own_pos", "own_hs", "own_vel", "own_vel_is_f32", "goal", "no_conflict",
                                          "ep_steps", "tick", "ipos", "ipos_is_f64", "ivel", "iflag", "ihs")]


class GcaMctsConfig(C.Structure):                   # gca_mcts_config
    _fields_ = [(n, C.c_double) for n in (
        "window_width", "window_height", "minimum_separation", "min_speed", "max_speed", "d_speed", "speed_sigma",
        "position_sigma", "d_heading", "heading_sigma")] + [("simulate_frame", C.c_int32), ("search_depth", C.c_int32),
                                                            ("random_intruders", C.c_int32), ("reserved0", C.c_int32),
                                                            ("turn_prob", C.c_double), ("turn_max_deg", C.c_double)]


def bench_workload_config():
    """gca_config of bench.py's workload (BASELINE.json configs[1]): SingleAircraft2Env with the package Config -
    the constants of PKG/config.py:6-39 and the reward row of PKG/SingleAircraft2Env.py:163-176, written out here so
    that the reference arm needs nothing of the product (tests/test_host_logic.py compares it field by field with
    variants.make_config("SingleAircraft2Env", Config))."""
    import math
    c = GcaConfig()
    c.window_width, c.window_height = 800.0, 800.0
    c.minimum_separation, c.nmac_dist, c.initial_min_dist, c.goal_radius = 555 / 30, 150 / 30, 3000 / 30, 600 / 30
    c.min_speed, c.max_speed, c.d_speed, c.speed_sigma = 50 / 30, 80 / 30, 5 / 30, 2 / 30
    c.d_heading, c.heading_sigma = math.radians(5), math.radians(2)
    c.ob_window_width, c.ob_window_height, c.ob_min_speed, c.ob_max_speed = 800.0, 800.0, 50 / 30, 80 / 30
    c.r_nmac, c.r_conflict, c.r_wall, c.r_goal, c.r_default = -5.0, -1.0, -100.0, 1.0, 0.0
    c.shaped_default, c.action_kind, c.obs_kind, c.wall_kind = 1, 1, 0, 1      # shaped, CONTINUOUS2, VECTOR, TERMINAL
    c.max_steps, c.time_limit, c.random_start, c.nearest_n = 0, 0, 0, 0
    c.ob_diagonal, c.conflict_coeff, c.goal_margin = 1.0, 0.0, 0.0
    c.shaped_nearest, c.intruder_turns = 0, 0
    c.position_drift, c.turn_prob, c.turn_max_deg = 0.0, 0.0, 0.0
    return c
