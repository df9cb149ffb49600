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


def make_input_reward_cfg(config):
    """gca_input_reward_cfg from a Simulators-style Config class (the attributes compute_input_reward reads,
    Simulators/SingleAircraftDiscrete9HEREnv.py:52-58, :244-276)."""
    c = config
    return abi.GcaInputRewardCfg(float(c.window_width), float(c.window_height), float(c.minimum_separation),
                                 float(c.NMAC_dist), float(c.goal_radius), float(c.conflict_penalty), float(c.NMAC_penalty),
                                 float(c.goal_reward), float(c.step_penalty), int(c.n), 1 if c.intruder_size != 0 else 0,
                                 1 if c.sparse_reward else 0, 0)


def compute_input_reward(new_inputs, config):
    """env.compute_input_reward (Simulators/SingleAircraftDiscrete9HEREnv.py:244-276) for a batch of relabelled
    (observation + goal) rows on the device: new_inputs CUDA tensor [M, dim], float32 or float64.
    Returns (reward float64 [M], done uint8 [M]) with done = r == 10 or r == -10 (agent_her.py:117)."""
    torch = _torch()
    lib = abi.load()
    x = new_inputs.contiguous()
    assert x.dim() == 2 and x.dtype in (torch.float32, torch.float64)
    m, dim = x.shape
    r = torch.empty((m,), dtype=torch.float64, device=x.device)
    d = torch.empty((m,), dtype=torch.uint8, device=x.device)
    cfg = make_input_reward_cfg(config)
    abi.check(lib.gca_input_reward(x.data_ptr(), m, dim, 1 if x.dtype == torch.float64 else 0, C.byref(cfg), r.data_ptr(),
                                   d.data_ptr(), x.device.index or 0,
                                   C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)))
    return r, d


def relabel_episode(obs, next_obs, goals, config, k=4, futures=None, generator=None):
    """The HER part of Agent.add (Algorithms/pytorch/agent_her.py:93-117) for one finished episode, on the device.

    obs, next_obs: CUDA tensors [T, dim_o] (s and s_n of every step), goals: [T, 2] (g).  For every step t and each of
    the k relabels, a future step f in [t, T) is drawn (np.random.randint(t, T); pass `futures` int64 [T, k] to replay
    recorded draws), its next observation's first two entries become the desired goal, and the reward of the relabelled
    transition is compute_input_reward(concat(new_ob, desired)).  Returns the dict of the (1 + k) T transitions the
    reference pushes into its memory, in its order: inputs, new_inputs [T, 1 + k, dim_o + 2], reward float64
    [T, 1 + k] (slot 0: NaN - the environment's reward of the original transition is the caller's), done uint8
    [T, 1 + k] (slot 0: 0 - likewise), futures [T, k]."""
    torch = _torch()
    T, dim_o = obs.shape
    dev = obs.device
    if futures is None:
        u = torch.rand((T, k), device=dev, generator=generator, dtype=torch.float64)
        t0 = torch.arange(T, device=dev, dtype=torch.float64)[:, None]
        futures = torch.minimum((t0 + torch.floor(u * (T - t0))).to(torch.int64), torch.full((1, 1), T - 1, device=dev))
    futures = futures.to(torch.int64)
    desired = next_obs[futures][..., :2]                                    # g_n[:2] of the future transition
    all_goals = torch.cat([goals[:, None, :].to(obs.dtype), desired], 1)    # slot 0: the episode's own goal
    ob_rep = obs[:, None, :].expand(T, 1 + k, dim_o)
    nob_rep = next_obs[:, None, :].expand(T, 1 + k, dim_o)
    inputs = torch.cat([ob_rep, all_goals], -1).contiguous()
    new_inputs = torch.cat([nob_rep, all_goals], -1).contiguous()
    r, d = compute_input_reward(new_inputs[:, 1:, :].reshape(T * k, dim_o + 2), config)
    reward = torch.full((T, 1 + k), float("nan"), dtype=torch.float64, device=dev)
    done = torch.zeros((T, 1 + k), dtype=torch.uint8, device=dev)
    reward[:, 1:] = r.view(T, k)
    done[:, 1:] = d.view(T, k)
    return {"inputs": inputs, "new_inputs": new_inputs, "reward": reward, "done": done, "futures": futures}


class AgentHerMemory(object):
    """The replay memory of the repo's own DQN-HER learner (Algorithms/pytorch/agent_her.py: `ReplayBuffer` :119-149, a
    deque(maxlen=buffer_size) of (state, action, reward, next_state, done), filled by `Agent.add` :93-117) as a ring of
    CUDA tensors.  `add_episode` is Agent.add for one finished episode - the transitions as they happened plus, with HER,
    k relabelled copies of each (future draws, desired goal, compute_input_reward: relabel_episode above), appended
    in the reference's order; `sample` returns what ReplayBuffer.sample returns (float states [batch, dim], long actions
    [batch, 1], float rewards / next_states / dones), drawn without replacement like random.sample."""

    def __init__(self, dim_inputs, buffer_size, batch_size, config, device=0, dtype=None, her=True, k=4, seed=0):
        torch = _torch()
        dev = torch.device("cuda", device)
        dt = dtype or torch.float32
        self.capacity, self.batch_size, self.config, self.her, self.k = int(buffer_size), int(batch_size), config, her, int(k)
        self.states = torch.zeros((self.capacity, dim_inputs), dtype=dt, device=dev)
        self.next_states = torch.zeros((self.capacity, dim_inputs), dtype=dt, device=dev)
        self.actions = torch.zeros((self.capacity, 1), dtype=torch.int64, device=dev)
        self.rewards = torch.zeros((self.capacity, 1), dtype=torch.float32, device=dev)
        self.dones = torch.zeros((self.capacity, 1), dtype=torch.float32, device=dev)
        self.size = 0
        self._next = 0
        self._gen = torch.Generator(device=dev)
        self._gen.manual_seed(int(seed))

    def __len__(self):
        return self.size

    def _append(self, states, actions, rewards, next_states, dones):
        torch = _torch()
        n = states.shape[0]
        idx = (self._next + torch.arange(n, device=states.device)) % self.capacity      # deque(maxlen): oldest out first
        self.states[idx] = states.to(self.states.dtype)
        self.next_states[idx] = next_states.to(self.states.dtype)
        self.actions[idx] = actions.reshape(n, 1).to(torch.int64)
        self.rewards[idx] = rewards.reshape(n, 1).to(torch.float32)
        self.dones[idx] = dones.reshape(n, 1).to(torch.float32)
        self._next = (self._next + n) % self.capacity
        self.size = min(self.capacity, self.size + n)

    def add_episode(self, obs, actions, rewards, next_obs, goals, dones, futures=None):
        """Agent.add(episode_experience, env): obs / next_obs [T, dim_o], actions [T], rewards [T], goals [T, 2],
        dones [T] as CUDA tensors (the episode's (s, a, r, s_n, g, done) tuples stacked)."""
        torch = _torch()
        T = obs.shape[0]
        if not self.her:
            self._append(torch.cat([obs, goals.to(obs.dtype)], -1), actions, rewards, torch.cat([next_obs, goals.to(obs.dtype)], -1), dones)
            return None
        tr = relabel_episode(obs, next_obs, goals, self.config, k=self.k, futures=futures, generator=self._gen)
        reward = tr["reward"].clone()
        done = tr["done"].to(torch.float32)
        reward[:, 0] = rewards.to(torch.float64)                     # slot 0: the transition as it happened
        done[:, 0] = dones.to(torch.float32)
        a = actions.reshape(T, 1).expand(T, 1 + self.k)
        d = tr["inputs"].shape[-1]
        self._append(tr["inputs"].reshape(-1, d), a.reshape(-1), reward.reshape(-1), tr["new_inputs"].reshape(-1, d),
                     done.reshape(-1))
        return tr

    def sample(self):
        torch = _torch()
        assert self.size >= self.batch_size
        idx = torch.randperm(self.size, device=self.states.device, generator=self._gen)[: self.batch_size]
        return (self.states[idx].float(), self.actions[idx], self.rewards[idx], self.next_states[idx].float(), self.dones[idx])
