"""HER replay on the device: baselines' episode buffer + "future" relabelling sampler.

Reference: Algorithms/baselines-master/baselines/her/replay_buffer.py (ReplayBuffer: store_episode, sample - which adds
o_2 = o[:, 1:], ag_2 = ag[:, 1:]) and her_sampler.py:19-61 (_sample_her_transitions).  Episodes produced by the batched
GoalEnv variants stay in HBM; `sample` is one kernel (csrc/gca_her.cu) that draws (episode, t), gathers the six rows of
each transition, substitutes a future achieved goal with probability 1 - 1 / (1 + replay_k) and recomputes the reward
with the env's compute_reward.  No CPU path.
"""
import ctypes as C

from . import abi


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise abi.GcaError("no CUDA device: the replay sampler has no CPU fallback")
    return torch


def sample_her_transitions(episode_batch, batch_size, replay_k, goal_radius, reward_kind, draws=None, seed=0, call=0,
                           return_draws=False):
    """_sample_her_transitions(episode_batch, batch_size_in_transitions) of her_sampler.py:19-61, strategy 'future'.

    episode_batch: dict of CUDA tensors o [E, T+1, dim_o], u [E, T, dim_u], g [E, T, 2], ag [E, T+1, 2], all float32 or
    all float64.  draws: optional dict of the sampler's four numpy draws (episode_idxs, t_samples int64; u_her,
    u_offset float64) as CUDA tensors - the replay-parity path; default Philox (seed, call).
    replay_k = 0 is strategy 'none' (future_p = 0).  Returns the transitions dict (o, u, g, ag, o_2, ag_2, r)."""
    torch = _torch()
    lib = abi.load()
    o, u, g, ag = (episode_batch[k].contiguous() for k in ("o", "u", "g", "ag"))
    assert o.dtype == u.dtype == g.dtype == ag.dtype and o.dtype in (torch.float32, torch.float64)
    E, T1, dim_o = o.shape
    T = T1 - 1
    assert u.shape[:2] == (E, T) and g.shape == (E, T, 2) and ag.shape == (E, T + 1, 2)
    dim_u = u.shape[2]
    dev, dt = o.device, o.dtype
    B = int(batch_size)
    out = {"o": torch.empty((B, dim_o), dtype=dt, device=dev), "u": torch.empty((B, dim_u), dtype=dt, device=dev),
           "g": torch.empty((B, 2), dtype=dt, device=dev), "ag": torch.empty((B, 2), dtype=dt, device=dev),
           "o_2": torch.empty((B, dim_o), dtype=dt, device=dev), "ag_2": torch.empty((B, 2), dtype=dt, device=dev),
           "r": torch.empty((B,), dtype=torch.float32, device=dev)}
    drawn = {k: torch.empty((B,), dtype=torch.int32, device=dev) for k in ("episode", "t", "future_t")}
    ep = abi.GcaHerEpisodes(o.data_ptr(), u.data_ptr(), g.data_ptr(), ag.data_ptr())
    tr = abi.GcaHerTransitions(*[out[k].data_ptr() for k in ("o", "u", "g", "ag", "o_2", "ag_2", "r")],
                               *[drawn[k].data_ptr() for k in ("episode", "t", "future_t")])
    dr = None
    if draws is not None:
        keep = [draws["episode_idxs"].to(torch.int64).contiguous(), draws["t_samples"].to(torch.int64).contiguous(),
                draws["u_her"].to(torch.float64).contiguous(), draws["u_offset"].to(torch.float64).contiguous()]
        dr = abi.GcaHerDraws(*[k.data_ptr() for k in keep])
    future_p = 1 - (1. / (1 + replay_k))                                  # her_sampler.py:14-17
    abi.check(lib.gca_her_sample(C.byref(ep), E, T, dim_o, dim_u, 2, 1 if dt == torch.float64 else 0, B, future_p,
                                 float(goal_radius), int(reward_kind), C.byref(dr) if dr is not None else None,
                                 int(seed), int(call), C.byref(tr), dev.index or 0,
                                 C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)))
    return (out, drawn) if return_draws else out


class HerReplayBuffer(object):
    """baselines' ReplayBuffer (replay_buffer.py) with the storage in HBM: `size_in_transitions` // T episodes of
    o [T+1, dim_o], u [T, dim_u], g [T, 2], ag [T+1, 2]; store_episode overwrites the oldest slots once full
    (the reference picks random slots then - `_get_storage_idx`; a ring keeps the device path free of host draws)."""

    def __init__(self, dim_o, dim_u, T, size_in_transitions, replay_k, goal_radius, reward_kind, device=0, dtype=None,
                 seed=0):
        torch = _torch()
        self.T, self.size = int(T), int(size_in_transitions) // int(T)
        self.replay_k, self.goal_radius, self.reward_kind, self.seed = replay_k, goal_radius, reward_kind, seed
        dt = dtype or torch.float32
        dev = torch.device("cuda", device)
        self.buffers = {"o": torch.zeros((self.size, T + 1, dim_o), dtype=dt, device=dev),
                        "u": torch.zeros((self.size, T, dim_u), dtype=dt, device=dev),
                        "g": torch.zeros((self.size, T, 2), dtype=dt, device=dev),
                        "ag": torch.zeros((self.size, T + 1, 2), dtype=dt, device=dev)}
        self.current_size = 0
        self.n_transitions_stored = 0
        self._next = 0
        self._calls = 0

    @property
    def full(self):
        return self.current_size == self.size

    def store_episode(self, episode_batch):
        """episode_batch: dict of tensors [rollout_batch_size, T or T+1, dim] (replay_buffer.py:62-77)."""
        n = episode_batch["u"].shape[0]
        torch = _torch()
        idx = (self._next + torch.arange(n, device=self.buffers["u"].device)) % self.size
        for k, buf in self.buffers.items():
            buf[idx] = episode_batch[k].to(buf.dtype)
        self._next = (self._next + n) % self.size
        self.current_size = min(self.size, self.current_size + n)
        self.n_transitions_stored += n * self.T

    def sample(self, batch_size):
        """replay_buffer.py:36-60: transitions of the episodes stored so far."""
        assert self.current_size > 0
        view = {k: v[: self.current_size] for k, v in self.buffers.items()}
        self._calls += 1
        return sample_her_transitions(view, batch_size, self.replay_k, self.goal_radius, self.reward_kind,
                                      seed=self.seed, call=self._calls)

    def get_current_episode_size(self):
        return self.current_size

    def get_current_size(self):
        return self.current_size * self.T

    def get_transitions_stored(self):
        return self.n_transitions_stored

    def clear_buffer(self):
        self.current_size = 0
