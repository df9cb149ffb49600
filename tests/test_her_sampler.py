"""HER replay sampler: the numpy restatement against outputs of the unmodified baselines
function (CPU), and the device kernel behind gca_her_sample against both (GPU)."""
import os

import numpy as np
import pytest

from helpers import GOLDEN
from oracle import her_sampler as ohs

CASES = ("her", "dher", "none")
KEYS = ("o", "u", "g", "ag", "o_2", "ag_2", "r")


def load(case):
    g = np.load(os.path.join(GOLDEN, "her_sampler.npz"))
    eb = {k: g["%s_ep_%s" % (case, k)] for k in ("o", "u", "g", "ag")}
    draws = {k: g["%s_draw_%s" % (case, k)] for k in ("episode_idxs", "t_samples", "u_her", "u_offset")}
    tr = {k: g["%s_tr_%s" % (case, k)] for k in KEYS}
    k, batch, radius, kind = g["%s_meta" % case]
    return eb, draws, tr, int(k), int(batch), float(radius), int(kind)


@pytest.mark.parametrize("case", CASES)
def test_oracle_equals_reference_sampler(case):
    eb, draws, want, k, batch, radius, kind = load(case)
    got, ft = ohs.sample_her_transitions(eb, batch, k, radius, kind, draws)
    for key in KEYS:
        assert got[key].shape == want[key].shape and np.array_equal(got[key], want[key]), key
    relabelled = int((ft >= 0).sum())
    assert (relabelled == 0) if k == 0 else (0.6 * batch < relabelled < 0.95 * batch)
    if case == "dher":
        assert 0 < want["r"].sum() < batch          # both reward values occur


@pytest.mark.gpu
@pytest.mark.parametrize("case", CASES)
def test_device_sampler_replays_reference(case):
    """f64 buffers + the recorded draws: every output bit equals what the reference function returned."""
    import torch
    from gca_b200 import replay
    eb, draws, want, k, batch, radius, kind = load(case)
    dev = {key: torch.as_tensor(v, device="cuda") for key, v in eb.items()}
    dd = {key: torch.as_tensor(v, device="cuda") for key, v in draws.items()}
    got, drawn = replay.sample_her_transitions(dev, batch, k, radius, kind, draws=dd, return_draws=True)
    for key in KEYS:
        assert np.array_equal(got[key].cpu().numpy(), want[key]), key
    _, ft = ohs.sample_her_transitions(eb, batch, k, radius, kind, draws)
    assert np.array_equal(drawn["future_t"].cpu().numpy(), ft)
    assert np.array_equal(drawn["episode"].cpu().numpy(), draws["episode_idxs"])


@pytest.mark.gpu
@pytest.mark.parametrize("dtype", ["float32", "float64"])
def test_device_sampler_philox_equals_oracle(dtype):
    """On-device Philox draws, ragged sizes, both dtypes, odd observation width (4-byte copies)."""
    import torch
    from gca_b200 import replay
    rng = np.random.RandomState(3)
    for E, T, dim_o, dim_u, batch, k, kind, radius in ((37, 50, 326, 2, 3001, 4, ohs.OBS_HER, 20.0),
                                                       (5, 9, 27, 1, 777, 4, ohs.OBS_DHER, 20.0),
                                                       (64, 20, 24, 1, 4096, 8, ohs.OBS_DHER, 20.0)):
        eb = {"o": rng.uniform(-1, 1, (E, T + 1, dim_o)), "u": rng.uniform(-1, 1, (E, T, dim_u)),
              "g": rng.uniform(100, 700, (E, T, 2)),
              "ag": np.cumsum(rng.normal(0, 7, (E, T + 1, 2)), 1) + rng.uniform(100, 700, (E, 1, 2))}
        eb = {key: v.astype(dtype) for key, v in eb.items()}
        draws = ohs.philox_draws(batch, E, T, seed=99, call=7)
        want, ft = ohs.sample_her_transitions(eb, batch, k, radius, kind, draws)
        if dtype == "float32":                      # the device compares the f32 norm with the f32 radius
            d = np.sqrt(((want["ag_2"] - want["g"]) ** 2).sum(-1, dtype=np.float32), dtype=np.float32)
            want["r"] = -(d > np.float32(radius)).astype(np.float32) if kind == ohs.OBS_HER else (d < np.float32(radius)).astype(np.float32)
        dev = {key: torch.as_tensor(v, device="cuda") for key, v in eb.items()}
        got, drawn = replay.sample_her_transitions(dev, batch, k, radius, kind, seed=99, call=7, return_draws=True)
        assert np.array_equal(drawn["episode"].cpu().numpy(), draws["episode_idxs"])
        assert np.array_equal(drawn["t"].cpu().numpy(), draws["t_samples"])
        assert np.array_equal(drawn["future_t"].cpu().numpy(), ft)
        for key in KEYS:
            assert np.array_equal(got[key].cpu().numpy(), want[key]), (key, E, T)


@pytest.mark.gpu
def test_replay_buffer_fed_from_the_batched_env():
    """Episodes of SingleAircraftHEREnv written straight from device observations, then sampled."""
    import torch
    from gca_b200 import abi, replay
    from gca_b200.batched import BatchedAircraftEnv
    from gym_guidance_collision_avoidance_single.envs.config import Config
    B, N, T = 64, 3, 16
    env = BatchedAircraftEnv("SingleAircraftHEREnv", B, Config, n_intruders=N, mode="fast", seed=1)
    buf = replay.HerReplayBuffer(dim_o=4 * N + 6, dim_u=2, T=T, size_in_transitions=4 * B * T, replay_k=4,
                                 goal_radius=Config.goal_radius, reward_kind=abi.OBS_HER)
    for rollout in range(3):
        env.reset()
        o, ag, g, u = [env.obs.clone()], [env.achieved.clone()], [], []
        for t in range(T):
            a = torch.rand((B, 2), device="cuda") * 2 - 1
            env.step(a, auto_reset=False)
            u.append(a); g.append(env.desired.clone()); o.append(env.obs.clone()); ag.append(env.achieved.clone())
        buf.store_episode({"o": torch.stack(o, 1), "u": torch.stack(u, 1), "g": torch.stack(g, 1), "ag": torch.stack(ag, 1)})
    assert buf.get_current_episode_size() == 3 * B and buf.get_transitions_stored() == 3 * B * T
    tr = buf.sample(1000)
    assert tr["o"].shape == (1000, 4 * N + 6) and tr["r"].shape == (1000,)
    assert torch.equal(tr["ag"], tr["o"][:, :2])             # achieved goal = normalised ownship position (Q13)
    assert torch.equal(tr["ag_2"], tr["o_2"][:, :2])
    assert float((tr["g"] != tr["g"][:1]).any()) == 1.0
    env.close()
