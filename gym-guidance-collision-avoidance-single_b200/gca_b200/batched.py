"""Device-resident batch of B independent SingleAircraft*Env instances.

This is the host side of the hot path: it owns the libgca handle, keeps the output tensors
on the GPU (PyTorch is only the allocator / stream provider) and exposes reset / step with
the reference's semantics for every env of the batch at once.  It collapses the reference's
L1 simulator (PKG/SingleAircraftEnv.py) and L3 VecEnv worker loop (baselines
dummy_vec_env.py:46-57) into one kernel launch per step.
"""
import ctypes as C

import numpy as np

from . import abi, variants

_STATE_FIELDS = (("own_pos", np.float32, (2,)), ("own_hs", np.float64, (2,)), ("own_vel", np.float64, (2,)),
                 ("own_vel_is_f32", np.uint8, ()), ("goal", np.float64, (2,)), ("no_conflict", np.int32, ()),
                 ("ep_steps", np.int32, ()), ("tick", np.uint32, ()),
                 ("ipos", np.float64, ("N", 2)), ("ipos_is_f64", np.uint8, ("N",)), ("ivel", np.float32, ("N", 2)),
                 ("iflag", np.uint8, ("N",)), ("ihs", np.float64, ("N", 2)))


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise abi.GcaError("no CUDA device: the batched simulator has no CPU fallback")
    return torch


class BatchedAircraftEnv(object):
    """B environments of one variant advanced by one fused CUDA kernel per step.

    variant   : a key of variants.VARIANTS ("SingleAircraftEnv", "SingleAircraft2Env", ...)
    config    : class with the reference's Config attributes (snapshotted now, like load_config)
    mode      : "fast" (f32 positions / observations) or "faithful" (reference mixed precision, f64 outputs)
    draws     : "philox" (on-device RNG) or "tape" (replay recorded numpy draws; parity tests)
    """

    def __init__(self, variant, num_envs, config, n_intruders=None, mode="fast", draws="philox", device=0, seed=0,
                 env_id0=0):
        torch = _torch()
        self.lib = abi.load()
        self.variant = variant
        self.config_class = config
        self.cfg = variants.make_config(variant, config)
        self.num_envs = int(num_envs)
        self.n_intruders = int(config.intruder_size if n_intruders is None else n_intruders)
        self.mode = {"fast": abi.MODE_FAST, "faithful": abi.MODE_FAITHFUL}[mode]
        self.draws = {"philox": abi.DRAWS_PHILOX, "tape": abi.DRAWS_TAPE}[draws]
        self.device = torch.device("cuda", device)
        self.real = torch.float64 if self.mode == abi.MODE_FAITHFUL else torch.float32
        self.obs_dim = variants.obs_dim(self.cfg, self.n_intruders)
        self.is_goal_env = self.cfg.obs_kind in (abi.OBS_HER, abi.OBS_DHER, abi.OBS_NEAREST)
        self.continuous = self.cfg.action_kind == abi.ACT_CONTINUOUS2
        handle = C.c_void_p()
        abi.check(self.lib.gca_create(C.byref(self.cfg), self.num_envs, self.n_intruders, self.mode, self.draws,
                                      device, int(seed) & (2 ** 64 - 1), int(env_id0), C.byref(handle)))
        self._h = handle
        B = self.num_envs
        with torch.cuda.device(self.device):
            self.obs = torch.zeros((B, max(self.obs_dim, 1)), dtype=self.real, device=self.device)
            self.achieved = torch.zeros((B, 2), dtype=self.real, device=self.device)
            self.desired = torch.zeros((B, 2), dtype=self.real, device=self.device)
            self.reward = torch.zeros((B,), dtype=self.real, device=self.device)
            self.done = torch.zeros((B,), dtype=torch.uint8, device=self.device)
            self.info = torch.zeros((B,), dtype=torch.uint8, device=self.device)
        # dist_nearest_intruder of the step (Simulators/SingleAircraftDiscrete3HEREnv.py:178), shaped_nearest variants only
        self.nearest = torch.zeros((B,), dtype=self.real, device=self.device) if self.cfg.shaped_nearest else None
        self._out = abi.GcaOut(self.obs.data_ptr() if self.obs_dim else None,
                               self.achieved.data_ptr() if self.is_goal_env else None,
                               self.desired.data_ptr() if self.is_goal_env else None,
                               self.reward.data_ptr(), self.done.data_ptr(), self.info.data_ptr(),
                               self.nearest.data_ptr() if self.nearest is not None else None)
        self._tape = None
        self._tape_keep = None
        self._host = None
        self.launches = 0

    # ------------------------------------------------------------------ lifetime
    def close(self):
        if getattr(self, "_h", None):
            self.lib.gca_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def set_tape(self, values, cursor=None):
        """Recorded draws for draws="tape": values [B, L] float64; cursor [B] int64 (default zeros)."""
        torch = _torch()
        v = torch.as_tensor(np.ascontiguousarray(values, np.float64), device=self.device)
        c = torch.zeros(self.num_envs, dtype=torch.int64, device=self.device) if cursor is None else \
            torch.as_tensor(np.ascontiguousarray(cursor, np.int64), device=self.device)
        self._tape_keep = (v, c)
        self._tape = abi.GcaTape(v.data_ptr(), v.shape[1], c.data_ptr())

    @property
    def tape_cursor(self):
        return self._tape_keep[1]

    def refresh_observation_params(self):
        """Re-read what _get_ob reads from the Config class on every call (Q12)."""
        variants.refresh_observation_params(self.cfg, self.config_class)
        abi.check(self.lib.gca_set_config(self._h, C.byref(self.cfg)))

    def _tape_ref(self):
        return C.byref(self._tape) if self._tape is not None else None

    # ------------------------------------------------------------------ device API
    def reset(self, mask=None):
        """reset() of every env (or those with mask != 0).  Returns the observation tensor [B, D]."""
        m = None
        if mask is not None:
            torch = _torch()
            m = torch.as_tensor(mask, device=self.device).to(torch.uint8).contiguous()
        abi.check(self.lib.gca_reset(self._h, m.data_ptr() if m is not None else None, self._tape_ref(),
                                     C.byref(self._out), self._stream()))
        self.launches += 1
        return self.obs

    def step(self, actions, auto_reset=True):
        """One step of all envs.  actions: device tensor, int32 [B] (discrete kinds) or real [B, 2]."""
        torch = _torch()
        if self.continuous:
            if actions.dtype != self.real or actions.shape != (self.num_envs, 2) or not actions.is_contiguous():
                actions = actions.to(self.real).reshape(self.num_envs, 2).contiguous()
        else:
            if actions.dtype != torch.int32 or actions.shape != (self.num_envs,) or not actions.is_contiguous():
                actions = actions.to(torch.int32).reshape(self.num_envs).contiguous()
        abi.check(self.lib.gca_step(self._h, actions.data_ptr(), self._tape_ref(), 1 if auto_reset else 0,
                                    C.byref(self._out), self._stream()))
        self.launches += self.kernels_per_step
        return self.obs, self.reward, self.done, self.info

    @property
    def kernels_per_step(self):
        """Kernels one step launches (gca_step_launches: 2 for a Philox handle with intruders)."""
        return self.lib.gca_step_launches(self._h)

    def check(self):
        """Synchronise and raise if an earlier asynchronous step failed on the device (gca_check)."""
        abi.check(self.lib.gca_check(self._h))

    def profile(self, on):
        """Per-kernel device timing of step() (CUDA events between the kernels; not inside a graph capture)."""
        abi.check(self.lib.gca_profile_enable(self._h, 1 if on else 0))

    def read_profile(self):
        """dict(steps, own_ms, intruders_ms, finish_ms, spawn_ms) summed over the steps recorded since the last read."""
        p = abi.GcaStepProfile()
        abi.check(self.lib.gca_profile_read(self._h, C.byref(p)))
        return {name: getattr(p, name) for name, _ in p._fields_}

    def observe(self):
        abi.check(self.lib.gca_observe(self._h, C.byref(self._out), self._stream()))
        self.launches += 1
        return self.obs

    def counters(self):
        """Device int32 [B, 4]: (no_conflict, steps of the current episode, Philox tick, finished episodes)."""
        torch = _torch()
        out = torch.empty((self.num_envs, 4), dtype=torch.int32, device=self.device)
        abi.check(self.lib.gca_read_counters(self._h, out.data_ptr(), self._stream()))
        return out

    # ------------------------------------------------------------------ host-buffer (end-to-end) API
    def _host_buffers(self, which=0):
        if self._host is None:
            self._host = [None, None]
        if self._host[which] is None:
            torch = _torch()
            B = self.num_envs
            pin = dict(pin_memory=True)
            h = {
                "actions": torch.zeros((B, 2), dtype=self.real, **pin) if self.continuous
                else torch.zeros((B,), dtype=torch.int32, **pin),
                "obs": torch.zeros((B, max(self.obs_dim, 1)), dtype=self.real, **pin),
                "achieved": torch.zeros((B, 2), dtype=self.real, **pin),
                "desired": torch.zeros((B, 2), dtype=self.real, **pin),
                "reward": torch.zeros((B,), dtype=self.real, **pin),
                "done": torch.zeros((B,), dtype=torch.uint8, **pin),
                "info": torch.zeros((B,), dtype=torch.uint8, **pin),
            }
            out = abi.GcaOut(h["obs"].data_ptr() if self.obs_dim else None,
                             h["achieved"].data_ptr() if self.is_goal_env else None,
                             h["desired"].data_ptr() if self.is_goal_env else None,
                             h["reward"].data_ptr(), h["done"].data_ptr(), h["info"].data_ptr())
            self._host[which] = (h, out, {k: v.numpy() for k, v in h.items()})
        return self._host[which]

    def host_io_bytes(self):
        """(host->device, device->host) bytes moved by one step_host call."""
        h, _, _ = self._host_buffers()
        nbytes = lambda t: t.numel() * t.element_size()
        d2h = nbytes(h["reward"]) + nbytes(h["done"]) + nbytes(h["info"])
        if self.obs_dim:
            d2h += nbytes(h["obs"])
        if self.is_goal_env:
            d2h += nbytes(h["achieved"]) + nbytes(h["desired"])
        return nbytes(h["actions"]), d2h

    def step_host(self, actions, auto_reset=True):
        """Step driven from host memory: numpy actions in, numpy (obs, reward, done, info) out.
        The returned arrays are views of pinned buffers that the next call overwrites."""
        h, out, views = self._host_buffers()
        views["actions"][...] = np.asarray(actions).reshape(views["actions"].shape)
        abi.check(self.lib.gca_step_host(self._h, h["actions"].data_ptr(), 1 if auto_reset else 0, C.byref(out)))
        self.launches += self.kernels_per_step
        self.last_host_views = views
        return views["obs"], views["reward"], views["done"], views["info"]

    def step_host_begin(self, actions, auto_reset=True):
        """VecEnv.step_async on host memory (gca_step_host_begin): returns at once; at most two steps in flight.  The
        download of this step overlaps the kernels of the next one begun before step_host_wait() is called."""
        which = self._host_turn = getattr(self, "_host_turn", 1) ^ 1
        h, out, views = self._host_buffers(which)
        views["actions"][...] = np.asarray(actions).reshape(views["actions"].shape)
        abi.check(self.lib.gca_step_host_begin(self._h, h["actions"].data_ptr(), 1 if auto_reset else 0, C.byref(out)))
        self.launches += self.kernels_per_step
        self._host_pending = getattr(self, "_host_pending", []) + [which]

    def step_host_wait(self):
        """VecEnv.step_wait: (obs, reward, done, info) of the oldest step begun - views of one of two pinned buffer sets,
        valid until the step after next is begun."""
        which = self._host_pending.pop(0)
        abi.check(self.lib.gca_step_host_wait(self._h))
        views = self.last_host_views = self._host_buffers(which)[2]
        return views["obs"], views["reward"], views["done"], views["info"]

    def reset_host(self):
        h, out, views = self._host_buffers()
        abi.check(self.lib.gca_reset_host(self._h, C.byref(out)))
        self.launches += 1
        self.last_host_views = views
        return views["obs"]

    # ------------------------------------------------------------------ full state
    def _state_arrays(self):
        B, N = self.num_envs, self.n_intruders
        return {name: np.zeros((B,) + tuple(N if s == "N" else s for s in shp), dt) for name, dt, shp in _STATE_FIELDS}

    def get_state(self):
        """Full simulator state as numpy arrays in the canonical layout of include/gca.h."""
        st = self._state_arrays()
        view = abi.GcaHostState(*[st[name].ctypes.data for name, _, _ in _STATE_FIELDS])
        abi.check(self.lib.gca_get_state(self._h, C.byref(view)))
        self.check()
        return st

    def set_state(self, state):
        keep = {}
        view = abi.GcaHostState()
        for name, dt, shp in _STATE_FIELDS:
            if name in state and state[name] is not None:
                a = np.ascontiguousarray(state[name], dt)
                want = (self.num_envs,) + tuple(self.n_intruders if s == "N" else s for s in shp)
                if a.shape != want:
                    raise ValueError("state[%r] has shape %r, expected %r" % (name, a.shape, want))
                keep[name] = a
                setattr(view, name, a.ctypes.data)
        abi.check(self.lib.gca_set_state(self._h, C.byref(view)))


def compute_reward(achieved_goal, desired_goal, radius, kind):
    """Batched HER relabel reward on device tensors (PKG/SingleAircraftHEREnv.py:194-196,
    PKG/SingleAircraftDiscreteHEREnv.py:184-186).  Returns float32 [...]."""
    torch = _torch()
    lib = abi.load()
    ag, g = achieved_goal, desired_goal
    if ag.dtype != g.dtype:                      # mixed f32/f64 promotes to f64, like numpy
        ag, g = ag.to(torch.float64), g.to(torch.float64)
    ag, g = ag.contiguous(), g.contiguous()
    m, n_ag = g.numel() // 2, ag.numel() // 2
    out = torch.empty(g.shape[:-1], dtype=torch.float32, device=ag.device)
    stream = C.c_void_p(torch.cuda.current_stream(ag.device).cuda_stream)
    is64 = 1 if ag.dtype == torch.float64 else 0
    if n_ag == m:
        abi.check(lib.gca_compute_reward(ag.data_ptr(), g.data_ptr(), m, float(radius), kind, is64, out.data_ptr(),
                                         ag.device.index or 0, stream))
    else:           # achieved [B, 2] against desired [k, B, 2]: the k relabels of every transition, ag not copied k times
        if n_ag <= 0 or m % n_ag:
            raise ValueError("desired_goal must hold a whole number of goals per achieved goal")
        abi.check(lib.gca_compute_reward_tiled(ag.data_ptr(), n_ag, g.data_ptr(), m, float(radius), kind, is64,
                                               out.data_ptr(), ag.device.index or 0, stream))
    return out
