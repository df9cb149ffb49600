"""ctypes wrapper of the CPU oracle (oracle/libgca_oracle.so).

TEST INFRASTRUCTURE, NOT PRODUCT: imported only by tests/, __graft_entry__.smoke() and the
cpu_baseline / --impl reference legs of bench.py.  It imports nothing of the product: its ctypes
struct definitions are its own (oracle/structs.py; a config built by the product's variants.make_config
has the same layout and is passed by address).
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(_HERE, "libgca_oracle.so")

TRIG_LIBM, TRIG_SHARED = 0, 1


def build(force=False):
    if force or not os.path.exists(LIB):
        subprocess.check_call(["make", "-C", _HERE] + (["-B"] if force else []))
    return LIB


def _abi():
    try:
        from . import structs
    except ImportError:                      # imported as a top-level module
        import structs
    return structs


class OracleBatch(C.Structure):
    pass


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    abi = _abi()
    build()
    OracleBatch._fields_ = [
        ("n_envs", C.c_int32), ("n_intr", C.c_int32),
        ("st", abi.GcaHostState),
        ("draws", C.c_int32), ("trig", C.c_int32),
        ("tape", C.c_void_p), ("tape_stride", C.c_int64), ("cursor", C.c_void_p),
        ("seed", C.c_uint64), ("env_id0", C.c_uint32), ("reserved0", C.c_uint32),
        ("f32_positions", C.c_int32), ("auto_reset", C.c_int32),
        ("obs", C.c_void_p), ("achieved", C.c_void_p), ("desired", C.c_void_p), ("reward", C.c_void_p),
        ("done", C.c_void_p), ("info", C.c_void_p), ("term_obs", C.c_void_p), ("nearest", C.c_void_p),
    ]
    L = C.CDLL(LIB)
    P = C.POINTER
    CFG = C.c_void_p                         # gca_config by address: the product's and the oracle's mirrors are the same bytes
    L.gca_oracle_step.argtypes = [CFG, P(OracleBatch), C.c_void_p]
    L.gca_oracle_reset.argtypes = [CFG, P(OracleBatch), C.c_void_p]
    L.gca_oracle_observe.argtypes = [CFG, P(OracleBatch)]
    L.gca_oracle_compute_reward.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p]
    L.gca_oracle_compute_reward_f32.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_double, C.c_int, C.c_void_p]
    L.gca_oracle_obs_dim.argtypes = [CFG, C.c_int]
    L.gca_oracle_philox4x32_10.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
    L.gca_oracle_philox4x32_10.restype = None
    L.gca_oracle_sincos.argtypes = [C.c_double, C.c_int, P(C.c_double), P(C.c_double)]
    L.gca_oracle_sincos.restype = None
    L.gca_oracle_log.argtypes = [C.c_double, C.c_int]
    L.gca_oracle_log.restype = C.c_double
    L.gca_oracle_philox_uniform2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
    L.gca_oracle_philox_uniform2.restype = None
    L.gca_oracle_philox_normal2.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.c_void_p]
    L.gca_oracle_philox_normal2.restype = None
    vp, i32, i64, u32, u64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_uint32, C.c_uint64, C.c_double
    MC = vp                                  # gca_mcts_config by address
    L.gca_oracle_mcts_move.argtypes = [MC, i32, vp, i32, i32, i32, vp, vp, u64, u32, i32, vp, vp]
    L.gca_oracle_mcts_rollout.argtypes = [MC, i32, vp, i32, i32, i32, vp, vp, u64, u32, u32, i32, vp, vp, vp]
    L.gca_oracle_mcts_playouts.argtypes = [MC, i32, vp, i64, i32, i32, vp, u64, u32, i32, vp, vp, vp]
    L.gca_oracle_mcts_search.argtypes = [MC, i32, vp, i32, i32, vp, vp, i32, vp, vp, vp, vp]
    L.gca_oracle_mcts_search_philox.argtypes = [MC, i32, vp, i64, i32, i32, u64, u32, i32, vp, vp, vp, vp]
    L.gca_oracle_raster.argtypes = [CFG, P(OracleBatch), vp, vp, vp]
    _lib = L
    return L


def _p(a):
    return None if a is None else a.ctypes.data


STATE_FIELDS = (("own_pos", np.float32, (2,)), ("own_hs", np.float64, (2,)), ("own_vel", np.float64, (2,)),
                ("own_vel_is_f32", np.uint8, ()), ("goal", np.float64, (2,)), ("no_conflict", np.int32, ()),
                ("ep_steps", np.int32, ()), ("tick", np.uint32, ()),
                ("ipos", np.float64, ("N", 2)), ("ipos_is_f64", np.uint8, ("N",)), ("ivel", np.float32, ("N", 2)),
                ("iflag", np.uint8, ("N",)), ("ihs", np.float64, ("N", 2)))


def empty_state(B, N):
    """Canonical host state (include/gca.h gca_host_state) as a dict of numpy arrays."""
    st = {}
    for name, dt, shp in STATE_FIELDS:
        shape = (B,) + tuple(N if s == "N" else s for s in shp)
        st[name] = np.zeros(shape, dt)
    return st


class OracleEnv(object):
    """B independent environments advanced sequentially on the CPU by the C restatement."""

    def __init__(self, cfg, n_envs, n_intr, draws=1, trig=TRIG_SHARED, seed=0, env_id0=0, f32_positions=False,
                 auto_reset=False, tape=None):
        self.L = lib()
        self.cfg = cfg
        self.B, self.N = int(n_envs), int(n_intr)
        self.D = self.L.gca_oracle_obs_dim(C.byref(cfg), self.N)
        self.state = empty_state(self.B, self.N)
        self.state["own_vel_is_f32"][...] = 1
        self.draws, self.trig, self.seed, self.env_id0 = draws, trig, seed, env_id0
        self.f32_positions, self.auto_reset = bool(f32_positions), bool(auto_reset)
        self.tape = None if tape is None else np.ascontiguousarray(tape, np.float64)
        self.cursor = np.zeros(self.B, np.int64)
        self.obs = np.zeros((self.B, self.D), np.float64)
        self.term_obs = np.zeros((self.B, self.D), np.float64)
        self.achieved = np.zeros((self.B, 2), np.float64)
        self.desired = np.zeros((self.B, 2), np.float64)
        self.reward = np.zeros(self.B, np.float64)
        self.done = np.zeros(self.B, np.uint8)
        self.info = np.zeros(self.B, np.uint8)
        self.nearest = np.zeros(self.B, np.float64)

    @property
    def own_vel_is_f32(self):
        return self.state["own_vel_is_f32"]

    def _batch(self):
        abi = _abi()
        b = OracleBatch()
        b.n_envs, b.n_intr = self.B, self.N
        st = abi.GcaHostState()
        for name, _, _ in STATE_FIELDS:
            if name == "ihs" and not self.cfg.intruder_turns:   # kept only by the variant whose intruders turn
                continue
            setattr(st, name, _p(self.state[name]))
        b.st = st
        b.draws, b.trig = self.draws, self.trig
        if self.tape is not None:
            b.tape, b.tape_stride = _p(self.tape), self.tape.shape[1]
        b.cursor = _p(self.cursor)
        b.seed, b.env_id0 = self.seed, self.env_id0
        b.f32_positions, b.auto_reset = int(self.f32_positions), int(self.auto_reset)
        b.obs, b.achieved, b.desired = _p(self.obs), _p(self.achieved), _p(self.desired)
        b.reward, b.done, b.info, b.term_obs = _p(self.reward), _p(self.done), _p(self.info), _p(self.term_obs)
        b.nearest = _p(self.nearest)
        return b

    def reset(self, mask=None):
        b = self._batch()
        m = None if mask is None else np.ascontiguousarray(mask, np.uint8)
        rc = self.L.gca_oracle_reset(C.byref(self.cfg), C.byref(b), _p(m))
        assert rc == 0, rc
        return self.obs

    def step(self, actions):
        a = np.zeros((self.B, 2), np.float64)
        actions = np.asarray(actions, np.float64)
        if actions.ndim == 1:
            a[:, 0] = actions
        else:
            a[:, : actions.shape[1]] = actions
        b = self._batch()
        rc = self.L.gca_oracle_step(C.byref(self.cfg), C.byref(b), _p(a))
        assert rc == 0, rc
        return self.obs, self.reward, self.done, self.info

    def raster(self, sprites, want_rgb=False):
        """Image observation of the current state: uint8 [B, H/4, W/4] (and the full RGB frames)."""
        W, H = int(self.cfg.window_width), int(self.cfg.window_height)
        sp = np.ascontiguousarray(sprites, np.uint8)
        frames = np.zeros((self.B, H // 4, W // 4), np.uint8)
        rgb = np.zeros((self.B, H, W, 3), np.uint8) if want_rgb else None
        b = self._batch()
        rc = self.L.gca_oracle_raster(C.byref(self.cfg), C.byref(b), _p(sp), _p(frames), _p(rgb))
        assert rc == 0, rc
        return (frames, rgb) if want_rgb else frames

    def observe(self):
        b = self._batch()
        rc = self.L.gca_oracle_observe(C.byref(self.cfg), C.byref(b))
        assert rc == 0, rc
        return self.obs


def compute_reward(ag, g, radius, kind):
    L = lib()
    ag = np.ascontiguousarray(ag)
    g = np.ascontiguousarray(g)
    m = ag.size // 2
    out = np.zeros(m, np.float32)
    if ag.dtype == np.float32 and g.dtype == np.float32:
        L.gca_oracle_compute_reward_f32(_p(ag), _p(g), m, float(radius), kind, _p(out))
    else:
        ag = np.ascontiguousarray(ag, np.float64)
        g = np.ascontiguousarray(g, np.float64)
        L.gca_oracle_compute_reward(_p(ag), _p(g), m, float(radius), kind, _p(out))
    return out.reshape(ag.shape[:-1])


# ------------------------------------------------------------------------------- MCTS model
def mcts_move(cfg, n, state, action, tape=None, trig=TRIG_LIBM, seed=0, root=0, first_frame=0):
    """SingleAircraftState.move on a copy of `state`; action = a0*3+a1.  Returns (state, flags, reward, draws used)."""
    L = lib()
    st = np.array(state, np.float64)
    flags = np.zeros(1, np.uint8)
    reward = np.zeros(1, np.float64)
    cur = np.zeros(1, np.int64)
    t = None if tape is None else np.ascontiguousarray(tape, np.float64)
    rc = L.gca_oracle_mcts_move(C.byref(cfg), n, _p(st), int(action), 0 if tape is not None else 1, trig, _p(t), _p(cur),
                                seed, root, first_frame, _p(flags), _p(reward))
    assert rc == 0
    return st, int(flags[0]), float(reward[0]), int(cur[0])


def mcts_rollout(cfg, n, root, depth, tape=None, trig=TRIG_LIBM, seed=0, root_id=0, playout=0, forced_first=-1):
    L = lib()
    r = np.ascontiguousarray(root, np.float64)
    reward = np.zeros(1, np.float64)
    first = np.zeros(1, np.int8)
    flags = np.zeros(1, np.uint8)
    cur = np.zeros(1, np.int64)
    t = None if tape is None else np.ascontiguousarray(tape, np.float64)
    rc = L.gca_oracle_mcts_rollout(C.byref(cfg), n, _p(r), depth, 0 if tape is not None else 1, trig, _p(t), _p(cur),
                                   seed, root_id, playout, forced_first, _p(reward), _p(first), _p(flags))
    assert rc == 0
    return float(reward[0]), int(first[0]), int(flags[0]), int(cur[0])


def mcts_playouts(cfg, n, roots, playouts, depth, first_action=None, seed=0, root_id0=0, trig=TRIG_SHARED):
    L = lib()
    roots = np.ascontiguousarray(roots, np.float64)
    R = roots.shape[0]
    rewards = np.zeros((R, playouts), np.float64)
    first = np.zeros((R, playouts), np.int8)
    flags = np.zeros((R, playouts), np.uint8)
    fa = None if first_action is None else np.ascontiguousarray(first_action, np.int8)
    rc = L.gca_oracle_mcts_playouts(C.byref(cfg), n, _p(roots), R, playouts, depth, _p(fa), seed, root_id0, trig,
                                    _p(rewards), _p(first), _p(flags))
    assert rc == 0
    return rewards, first, flags


def mcts_search(cfg, n, root, sims, depth, tape, trig=TRIG_LIBM):
    L = lib()
    r = np.ascontiguousarray(root, np.float64)
    t = np.ascontiguousarray(tape, np.float64)
    cur = np.zeros(1, np.int64)
    best = C.c_int(-1)
    cn, cq, ca = np.zeros(9), np.zeros(9), np.zeros(9, np.int32)
    rc = L.gca_oracle_mcts_search(C.byref(cfg), n, _p(r), sims, depth, _p(t), _p(cur), trig, C.byref(best), _p(cn),
                                  _p(cq), _p(ca))
    assert rc == 0
    return best.value, cn, cq, ca, int(cur[0])


def mcts_search_philox(cfg, n, roots, sims, depth, seed=0, root_id0=0, trig=TRIG_SHARED):
    """The contract of gca_mcts_search on host arrays: (best_action int32 [R], child_n, child_q f64 [R, 9],
    child_action int32 [R, 9])."""
    L = lib()
    roots = np.ascontiguousarray(roots, np.float64)
    R = roots.shape[0]
    best = np.zeros(R, np.int32)
    cn, cq, ca = np.zeros((R, 9)), np.zeros((R, 9)), np.zeros((R, 9), np.int32)
    rc = L.gca_oracle_mcts_search_philox(C.byref(cfg), n, _p(roots), R, sims, depth, seed, root_id0, trig, _p(best),
                                         _p(cn), _p(cq), _p(ca))
    assert rc == 0
    return best, cn, cq, ca
