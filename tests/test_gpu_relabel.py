"""Device relabel of the repo's own HER learner (remainder): gca_input_reward against the values
recorded from the reference's compute_input_reward (tests/golden/d9her_reward.npz, made by make_golden.py from the
unmodified Simulators/SingleAircraftDiscrete9HEREnv.py:244-276), and relabel_episode against a line-by-line restatement
of Algorithms/pytorch/agent_her.py:93-117 driven by the same future draws."""
import math
import os

import numpy as np
import pytest

from helpers import GOLDEN

pytestmark = pytest.mark.gpu


def _config():
    from Simulators.config import Config

    class C(Config):
        intruder_size = 8
    return C


def _input_reward_numpy(v, c):
    """compute_input_reward with NumPy scalars of v's dtype, statement by statement (:244-276); squares as products."""
    W, H = v.dtype.type(c.window_width), v.dtype.type(c.window_height)

    def metric(x1, y1, x2, y2):
        dx, dy = x1 - x2, y1 - y2
        return math.sqrt(dx * dx + dy * dy)
    ownx, owny, gx, gy = v[0] * W, v[1] * H, v[-2] * W, v[-1] * H
    dg = metric(ownx, owny, gx, gy)
    if c.intruder_size != 0:
        for idx in range(c.n):
            d = metric(ownx, owny, v[idx * 4 + 4] * W, v[idx * 4 + 5] * H)
            if d < c.minimum_separation:
                return c.NMAC_penalty if d < c.NMAC_dist else c.conflict_penalty
    if dg < c.goal_radius:
        return c.goal_reward
    return c.step_penalty if c.sparse_reward else -dg / 1200


def test_input_reward_matches_reference_golden():
    import torch
    from gca_b200 import replay
    g = np.load(os.path.join(GOLDEN, "d9her_reward.npz"))
    c = _config()
    r, d = replay.compute_input_reward(torch.as_tensor(g["inputs"], device="cuda"), c)
    r, want = r.cpu().numpy(), g["ri"]
    # the branch taken (conflict / NMAC / goal / default) is exact; the shaped default -dist / 1200 is the reference's
    # value to 1e-15 relative (the reference squares with pow(x, 2), this kernel with x * x: at most an ulp apart)
    kinds = lambda x: np.select([x == c.NMAC_penalty, x == c.conflict_penalty, x == c.goal_reward], [1, 2, 3], 0)
    assert np.array_equal(kinds(r), kinds(want))
    assert len(set(kinds(want).tolist())) == 4                          # every branch occurs in the fixture
    assert np.allclose(r, want, rtol=1e-15, atol=0)
    assert np.array_equal(d.cpu().numpy(), ((want == 10) | (want == -10)).astype(np.uint8))
    # f32 rows (what the learner concatenates from the float32 observations): NumPy-scalar arithmetic in f32
    x32 = g["inputs"].astype(np.float32)
    r32, _ = replay.compute_input_reward(torch.as_tensor(x32, device="cuda"), c)
    want32 = np.array([_input_reward_numpy(v, c) for v in x32])
    assert np.array_equal(r32.cpu().numpy(), want32)


def test_relabel_episode_equals_agent_her_add():
    """Agent.add's HER branch (agent_her.py:103-117) for one episode: same futures -> same inputs, rewards, dones."""
    import torch
    from gca_b200 import replay
    c = _config()
    rng = np.random.RandomState(3)
    T, dim_o, k = 37, 24, 4
    s = rng.uniform(0.05, 0.95, (T, dim_o)).astype(np.float32)
    s_n = np.roll(s, -1, axis=0) + rng.normal(0, 0.004, (T, dim_o)).astype(np.float32)
    s_n[::5, 4:6] = s_n[::5, 0:2] + rng.uniform(-0.02, 0.02, (len(s_n[::5]), 2)).astype(np.float32)   # conflicts
    goal = np.repeat(rng.uniform(0.1, 0.9, (1, 2)).astype(np.float32), T, 0)
    futures = np.stack([rng.randint(t, T, k) for t in range(T)])                    # np.random.randint(t, len(episode))
    out = replay.relabel_episode(torch.as_tensor(s, device="cuda"), torch.as_tensor(s_n, device="cuda"),
                                 torch.as_tensor(goal, device="cuda"), c, k=k,
                                 futures=torch.as_tensor(futures, device="cuda"))
    inputs, new_inputs = out["inputs"].cpu().numpy(), out["new_inputs"].cpu().numpy()
    reward, done = out["reward"].cpu().numpy(), out["done"].cpu().numpy()
    for t in range(T):
        assert np.array_equal(inputs[t, 0], np.concatenate([s[t], goal[t]]))
        assert np.array_equal(new_inputs[t, 0], np.concatenate([s_n[t], goal[t]]))
        assert np.isnan(reward[t, 0]) and done[t, 0] == 0
        for j in range(k):
            desired = s_n[futures[t, j]][:2]
            ni = np.concatenate([s_n[t], desired])
            assert np.array_equal(inputs[t, 1 + j], np.concatenate([s[t], desired]))
            assert np.array_equal(new_inputs[t, 1 + j], ni)
            r_n = _input_reward_numpy(ni, c)
            assert reward[t, 1 + j] == r_n
            assert done[t, 1 + j] == (1 if (r_n == 10 or r_n == -10) else 0)
    assert (reward[:, 1:] == c.conflict_penalty).any() and (reward[:, 1:] == c.goal_reward).any()
    # device-drawn futures: every draw lies in [t, T)
    out2 = replay.relabel_episode(torch.as_tensor(s, device="cuda"), torch.as_tensor(s_n, device="cuda"),
                                  torch.as_tensor(goal, device="cuda"), c, k=k)
    f2 = out2["futures"].cpu().numpy()
    assert (f2 >= np.arange(T)[:, None]).all() and (f2 < T).all()


def test_agent_her_memory_holds_what_agent_add_pushes():
    """AgentHerMemory.add_episode = Agent.add (agent_her.py:93-117): (1 + k) T entries in the reference's order
    (the transition as it happened, then its k relabels), ring semantics of deque(maxlen), sample() shapes / dtypes."""
    import torch
    from gca_b200 import replay
    c = _config()
    rng = np.random.RandomState(8)
    T, dim_o, k = 11, 24, 4
    s = rng.uniform(0.05, 0.95, (T, dim_o)).astype(np.float32)
    s_n = np.roll(s, -1, axis=0)
    goal = np.repeat(rng.uniform(0.1, 0.9, (1, 2)).astype(np.float32), T, 0)
    a = rng.randint(0, 9, T)
    r = rng.uniform(-1, 0, T)
    dn = np.zeros(T); dn[-1] = 1
    futures = np.stack([rng.randint(t, T, k) for t in range(T)])
    mem = replay.AgentHerMemory(dim_o + 2, buffer_size=3 * T * (1 + k) - 7, batch_size=32, config=c, k=k, seed=1)
    dev = lambda x: torch.as_tensor(x, device="cuda")
    for rep in range(3):                                             # the third episode wraps the ring
        mem.add_episode(dev(s), dev(a), dev(r), dev(s_n), dev(goal), dev(dn), futures=dev(futures))
    assert len(mem) == mem.capacity
    # the last episode's entries are the newest ones of the ring, in Agent.add's order
    n = T * (1 + k)
    idx = (mem._next - n + np.arange(n)) % mem.capacity
    st, ns = mem.states.cpu().numpy()[idx], mem.next_states.cpu().numpy()[idx]
    rw, dd, ac = mem.rewards.cpu().numpy()[idx, 0], mem.dones.cpu().numpy()[idx, 0], mem.actions.cpu().numpy()[idx, 0]
    j = 0
    for t in range(T):
        assert np.array_equal(st[j], np.concatenate([s[t], goal[t]])) and np.array_equal(ns[j], np.concatenate([s_n[t], goal[t]]))
        assert rw[j] == np.float32(r[t]) and dd[j] == dn[t] and ac[j] == a[t]
        j += 1
        for q in range(k):
            desired = s_n[futures[t, q]][:2]
            assert np.array_equal(st[j], np.concatenate([s[t], desired]))
            assert rw[j] == np.float32(_input_reward_numpy(np.concatenate([s_n[t], desired]), c)) and ac[j] == a[t]
            j += 1
    states, actions, rewards, next_states, dones = mem.sample()
    assert states.shape == (32, dim_o + 2) and states.dtype == torch.float32 and actions.dtype == torch.int64
    assert actions.shape == rewards.shape == dones.shape == (32, 1) and next_states.shape == states.shape
