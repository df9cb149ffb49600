"""ctypes mirror of include/gca.h and the loader of libgca.so.

There is no CPU implementation behind this module: if the shared library has not been built
(`python __graft_entry__.py build`) or no CUDA device is usable, the calls raise.
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("GCA_LIB") or os.path.join(os.path.dirname(_HERE), "lib", "libgca.so")

GCA_ABI_VERSION = 7

MODE_FAITHFUL, MODE_FAST = 0, 1
DRAWS_TAPE, DRAWS_PHILOX = 0, 1
ACT_DISCRETE9, ACT_CONTINUOUS2, ACT_DISCRETE3, ACT_DISCRETE3_HEADING = 0, 1, 2, 3
OBS_VECTOR, OBS_HER, OBS_DHER, OBS_RAW, OBS_NONE, OBS_NEAREST, OBS_RAW6 = 0, 1, 2, 3, 4, 5, 6
WALL_NONE, WALL_TERMINAL, WALL_PENALTY = 0, 1, 2
INFO_NONE, INFO_NMAC, INFO_CONFLICT, INFO_GOAL, INFO_WALL, INFO_MAXSTEPS = range(6)
INFO_STR = ("", "n", "c", "g", "w", "m")

SLOT_OWNSHIP, SLOT_GOAL, SLOT_RESET, SLOT_TURN = 0x80000000, 0x40000000, 0x20000000, 0x08000000


class GcaConfig(C.Structure):
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


class GcaHostState(C.Structure):
    _fields_ = [("own_pos", C.c_void_p), ("own_hs", C.c_void_p), ("own_vel", C.c_void_p),
                ("own_vel_is_f32", C.c_void_p), ("goal", C.c_void_p),
                ("no_conflict", C.c_void_p), ("ep_steps", C.c_void_p), ("tick", C.c_void_p), ("ipos", C.c_void_p),
                ("ipos_is_f64", C.c_void_p), ("ivel", C.c_void_p), ("iflag", C.c_void_p), ("ihs", C.c_void_p)]


class GcaOut(C.Structure):
    _fields_ = [("obs", C.c_void_p), ("achieved", C.c_void_p), ("desired", C.c_void_p), ("reward", C.c_void_p),
                ("done", C.c_void_p), ("info", C.c_void_p), ("nearest", C.c_void_p)]


class GcaTape(C.Structure):
    _fields_ = [("values", C.c_void_p), ("stride", C.c_int64), ("cursor", C.c_void_p)]


class GcaHerEpisodes(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("o", "u", "g", "ag")]


class GcaHerDraws(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("episode_idxs", "t_samples", "u_her", "u_offset")]


class GcaHerTransitions(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("o", "u", "g", "ag", "o_2", "ag_2", "r", "episode", "t", "future_t")]


class GcaInputRewardCfg(C.Structure):
    """gca_input_reward_cfg of include/gca.h"""
    _fields_ = [(n, C.c_double) for n in ("window_width", "window_height", "minimum_separation", "nmac_dist", "goal_radius",
                                          "conflict_penalty", "nmac_penalty", "goal_reward", "step_penalty")] + \
               [(n, C.c_int32) for n in ("n_listed", "has_intruders", "sparse_reward", "reserved")]


class GcaMctsConfig(C.Structure):
    _fields_ = [(n, C.c_double) for n in (
        "window_width", "window_height", "minimum_separation", "min_speed", "max_speed", "d_speed", "speed_sigma",
        "position_sigma", "d_heading", "heading_sigma")] + [("simulate_frame", C.c_int32), ("search_depth", C.c_int32),
                                                            ("random_intruders", C.c_int32), ("reserved0", C.c_int32),
                                                            ("turn_prob", C.c_double), ("turn_max_deg", C.c_double)]


MCTS_WALL, MCTS_CONFLICT, MCTS_GOAL = 1, 2, 4
STAT_NAMES = ("steps", "episodes", "nmac", "conflict_steps", "goal", "wall", "maxsteps")


def make_mcts_config(cfg_cls, random_intruders=False):
    """Snapshot Algorithms/MCTS/config_single.py-style class attributes.  random_intruders selects the model of
    Algorithms/MCTS/nodes_single_randintru.py (6 N + 8 state vectors, intruders that turn with probability 0.1 per
    sub-frame by up to 10 degrees, :64-65)."""
    c = GcaMctsConfig()
    for name in ("window_width", "window_height", "minimum_separation", "min_speed", "max_speed", "d_speed",
                 "speed_sigma", "position_sigma", "d_heading", "heading_sigma"):
        setattr(c, name, float(getattr(cfg_cls, name)))
    c.simulate_frame = int(cfg_cls.simulate_frame)
    c.search_depth = int(cfg_cls.search_depth)
    c.random_intruders = 1 if random_intruders else 0
    c.turn_prob = 0.1 if random_intruders else 0.0
    c.turn_max_deg = 10.0 if random_intruders else 0.0
    return c


class GcaStepProfile(C.Structure):
    """gca_step_profile of include/gca.h"""
    _fields_ = [("steps", C.c_int64), ("own_ms", C.c_double), ("intruders_ms", C.c_double), ("finish_ms", C.c_double),
                ("spawn_ms", C.c_double)]


class GcaError(RuntimeError):
    pass


_lib = None


def load():
    """Load libgca.so (once).  Fails loudly: there is no fallback path."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise GcaError("libgca.so is not built (%s missing); run `python __graft_entry__.py build`. "
                       "There is no CPU fallback." % LIB_PATH)
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, u32, u64 = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64
    P = C.POINTER
    sigs = {
        "gca_abi_version": ([], C.c_int),
        "gca_last_error": ([], C.c_char_p),
        "gca_obs_dim": ([P(GcaConfig), i32], C.c_int),
        "gca_create": ([P(GcaConfig), i32, i32, i32, i32, i32, u64, u32, P(vp)], C.c_int),
        "gca_destroy": ([vp], C.c_int),
        "gca_set_config": ([vp, P(GcaConfig)], C.c_int),
        "gca_set_seed": ([vp, u64], C.c_int),
        "gca_reset": ([vp, vp, P(GcaTape), P(GcaOut), vp], C.c_int),
        "gca_step": ([vp, vp, P(GcaTape), i32, P(GcaOut), vp], C.c_int),
        "gca_step_host": ([vp, vp, i32, P(GcaOut)], C.c_int),
        "gca_reset_host": ([vp, P(GcaOut)], C.c_int),
        "gca_step_host_begin": ([vp, vp, i32, P(GcaOut)], C.c_int),
        "gca_step_host_wait": ([vp], C.c_int),
        "gca_get_state": ([vp, P(GcaHostState)], C.c_int),
        "gca_set_state": ([vp, P(GcaHostState)], C.c_int),
        "gca_observe": ([vp, P(GcaOut), vp], C.c_int),
        "gca_read_counters": ([vp, vp, vp], C.c_int),
        "gca_step_launches": ([vp], C.c_int),
        "gca_check": ([vp], C.c_int),
        "gca_profile_enable": ([vp, i32], C.c_int),
        "gca_profile_read": ([vp, P(GcaStepProfile)], C.c_int),
        "gca_compute_reward": ([vp, vp, i64, C.c_double, i32, i32, vp, i32, vp], C.c_int),
        "gca_compute_reward_tiled": ([vp, i64, vp, i64, C.c_double, i32, i32, vp, i32, vp], C.c_int),
        "gca_input_reward": ([vp, i64, i32, i32, P(GcaInputRewardCfg), vp, vp, i32, vp], C.c_int),
        "gca_raster": ([vp, vp, vp, i64, i64, i32, i32, vp, vp], C.c_int),
        "gca_monitor_update": ([vp, i32, vp, i64, vp, vp, vp, i64, vp, u32, i32, vp], C.c_int),
        "gca_stats_update": ([vp, vp, i64, vp, i32, vp], C.c_int),
        "gca_her_sample": ([P(GcaHerEpisodes), i64, i32, i32, i32, i32, i32, i64, C.c_double, C.c_double, i32,
                            P(GcaHerDraws), u64, u32, P(GcaHerTransitions), i32, vp], C.c_int),
        "gca_mcts_move": ([P(GcaMctsConfig), i32, vp, vp, vp, i64, P(GcaTape), u64, u32, i32, i32, vp], C.c_int),
        "gca_mcts_playouts": ([P(GcaMctsConfig), i32, vp, i64, i32, i32, vp, u64, u32, vp, vp, vp, i32, vp], C.c_int),
        "gca_mcts_search_workspace": ([P(GcaMctsConfig), i32, i64, i32, i32], C.c_int64),
        "gca_mcts_search": ([P(GcaMctsConfig), i32, vp, i64, i32, i32, u64, u32, vp, i64, vp, vp, vp, vp, i32, vp], C.c_int),
    }
    for name, (argtypes, restype) in sigs.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = restype
    if lib.gca_abi_version() != GCA_ABI_VERSION:
        raise GcaError("libgca.so ABI version %d != expected %d" % (lib.gca_abi_version(), GCA_ABI_VERSION))
    _lib = lib
    return lib


def check(status):
    if status != 0:
        msg = load().gca_last_error()
        raise GcaError("libgca call failed (%d): %s" % (status, msg.decode() if msg else "?"))
