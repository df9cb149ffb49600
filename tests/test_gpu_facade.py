"""The reference-facing gym facade (B = 1) and the VecEnv, replayed against the reference traces
through their public API: same call sequence a user of the reference would write."""
import numpy as np
import pytest

from helpers import config_class, load_trace

pytestmark = pytest.mark.gpu
RTOL = 1e-9


def close(a, b):
    return np.allclose(a, b, rtol=RTOL, atol=1e-300)


def _make(vk, n, **kw):
    cfgc = config_class(vk)
    cfgc.intruder_size = n                      # the reference idiom: mutate Config, then construct
    try:
        if vk == "mcts":
            from Simulators.SingleAircraftMCTSEnv import SingleAircraftEnv as cls
        elif vk == "mctsrnd":
            from Simulators.SingleAircraftMCTSRandIntruderEnv import SingleAircraftEnv as cls
        elif vk == "d9her":
            from Simulators.SingleAircraftDiscrete9HEREnv import SingleAircraftDiscrete9HEREnv as cls
        elif vk == "d3her":
            from Simulators.SingleAircraftDiscrete3HEREnv import SingleAircraftDiscrete3HEREnv as cls
        elif vk == "simenv":
            from Simulators.SingleAircraftEnv import SingleAircraftEnv as cls
        elif vk == "rndenv":
            from Simulators.SingleAircraftRandomEnv import SingleAircraftRandomEnv as cls
        else:
            import gym_guidance_collision_avoidance_single.envs as envs
            cls = {"env": envs.SingleAircraftEnv, "env2": envs.SingleAircraft2Env, "her": envs.SingleAircraftHEREnv,
                   "dher": envs.SingleAircraftDiscreteHEREnv}[vk]
        return cls(**kw)
    finally:
        cfgc.intruder_size = 80 if vk in ("mcts", "mctsrnd", "d9her", "d3her", "simenv", "rndenv") else 0


def _ref_action(vk, a):
    if vk in ("env", "dher", "d9her", "d3her", "simenv", "rndenv"):
        return int(a[0])
    if vk in ("mcts", "mctsrnd"):
        return (int(a[0]), int(a[1]))
    return np.array(a, np.float64)


@pytest.mark.parametrize("vk,n", [("env", 3), ("env", 80), ("env2", 3), ("her", 3), ("dher", 3), ("mcts", 80),
                                  ("d9her", 12), ("d9her", 80), ("d3her", 12), ("simenv", 3), ("rndenv", 80),
                                  ("mctsrnd", 3), ("mctsrnd", 80)])
def test_single_env_api_replays_reference_trace(vk, n):
    g = load_trace(vk, n)
    plain = [int(i) for i in np.nonzero(g["kind_id"] == 0)[0]][:2]
    for tr in plain:
        tape = np.nan_to_num(g["tape"][tr:tr + 1], nan=0.0)
        if vk in ("her", "dher", "d9her", "d3her"):   # these constructors reset() once themselves (PKG/SingleAircraftHEREnv.py:32)
            tape = np.concatenate([tape[:, : int(g["cur_reset0"][tr])], tape], axis=1)
        env = _make(vk, n, draws="tape", tape=tape)
        ob = env.reset()
        her = isinstance(ob, dict)
        assert close(ob["observation"] if her else ob, g["obs0"][tr])
        if her:
            assert ob["achieved_goal"].dtype == np.float32 and ob["desired_goal"].dtype == np.float64
            assert close(ob["achieved_goal"], g["ag0"][tr]) and close(ob["desired_goal"], g["dg0"][tr])
        else:
            assert ob.dtype == np.float64 and ob.shape == ((6 if vk == "mctsrnd" else 4) * n + 8,)
        for t in range(g["actions"].shape[1]):
            ob, r, done, info = env.step(_ref_action(vk, g["actions"][tr, t]))
            assert close(ob["observation"] if her else ob, g["obs"][tr, t])
            assert close(r, g["reward"][tr, t])
            if vk not in ("mcts", "mctsrnd", "d9her", "d3her", "simenv", "rndenv"):
                assert isinstance(r, int) == bool(g["reward_is_int"][tr, t]), (vk, t, r)
            assert done == bool(g["done"][tr, t]) and isinstance(done, bool)
            code = ("", "n", "c", "g", "w", "m")[g["event"][tr, t]]
            if vk in ("env", "env2", "mctsrnd"):     # (the random-intruder env returns the bare string, :164)
                assert info == code
            elif vk == "dher":
                assert info == {}
            elif vk == "d3her":
                assert close(info, g["nearest"][tr, t])
            else:
                assert info == {"result": code}
            assert env.no_conflict == int(g["no_conflict"][tr, t])
            if done:
                ob = env.reset()
                assert close(ob["observation"] if her else ob, g["reset_obs"][tr, t])
        env.close()


def test_spaces_and_attributes_match_reference():
    env = _make("env", 5)
    assert env.observation_space.shape == (28,) and env.observation_space.dtype == np.float32
    assert env.action_space.n == 9 and env.intruder_size == 5
    assert env.seed(3) == [3]
    assert env.minimum_separation == 18.5 and env.NMAC_dist == 5.0 and env.goal_radius == 20.0
    env.close()
    e2 = _make("env2", 0)
    assert e2.action_space.shape == (2,) and e2.observation_space.shape == (8,)
    e2.reset()
    with pytest.raises(AssertionError):
        e2.step(np.array([1.5, 0.0]))                # PKG/SingleAircraft2Env.py:127
    e2.close()
    her = _make("her", 2)
    sp = her.observation_space.spaces
    assert sp["observation"].shape == (14,) and sp["achieved_goal"].shape == (2,) and sp["desired_goal"].shape == (2,)
    r = her.compute_reward(np.array([0.1, 0.2]), np.array([0.7, 0.9]), None)
    assert r == 0.0 and np.signbit(r) and r.dtype == np.float32            # always -0.0 (Q14)
    rb = her.compute_reward(np.zeros((5, 4, 2)), np.ones((5, 4, 2)), None)
    assert rb.shape == (5, 4)
    her.close()
    dher = _make("dher", 2)
    assert dher.action_space.n == 3
    assert dher.compute_reward(np.array([100.0, 100.0]), np.array([110.0, 100.0]), None) == 1.0
    assert dher.compute_reward(np.array([100.0, 100.0]), np.array([130.0, 100.0]), None) == 0.0
    dher.close()


def test_registered_ids_and_time_limit():
    import gym_guidance_collision_avoidance_single as pkg
    assert set(pkg.registry) >= {"guidance-collision-avoidance-single-v0",
                                 "guidance-collision-avoidance-single-continuous-action-v0"}
    env = pkg.make("guidance-collision-avoidance-single-v0", time_limit=7)   # spec default is 10000
    env.reset()
    dones = [env.step(4)[2] for _ in range(7)]
    assert dones == [False] * 6 + [True]
    env.close()


def test_vec_env_interface():
    import torch
    from gca_b200.vec_env import AircraftVecEnv, AlreadySteppingError, NotSteppingError
    venv = AircraftVecEnv("guidance-collision-avoidance-single-v0", 512, n_intruders=20, seed=1)
    assert venv.num_envs == 512 and venv.observation_space.shape == (88,) and venv.action_space.n == 9
    obs = venv.reset()
    assert obs.shape == (512, 88) and obs.dtype == torch.float32 and obs.is_cuda
    with pytest.raises(NotSteppingError):
        venv.step_wait()
    a = torch.randint(0, 9, (512,), device="cuda", dtype=torch.int32)
    venv.step_async(a)
    with pytest.raises(AlreadySteppingError):
        venv.step_async(a)
    obs, rew, done, info = venv.step_wait()
    assert rew.shape == (512,) and done.shape == (512,) and info.shape == (512,)
    # host flavour: numpy in, numpy out, same numbers as the device flavour
    hv = AircraftVecEnv("guidance-collision-avoidance-single-v0", 512, n_intruders=20, seed=1, host=True)
    venv2 = AircraftVecEnv("guidance-collision-avoidance-single-v0", 512, n_intruders=20, seed=1)
    ho = hv.reset()
    assert isinstance(ho, np.ndarray) and np.array_equal(ho, venv2.reset().cpu().numpy())
    for _ in range(5):
        an = np.random.randint(0, 9, 512).astype(np.int32)
        o_h, r_h, d_h, i_h = hv.step(an)
        o_d, r_d, d_d, i_d = venv2.step(torch.as_tensor(an, device="cuda"))
        assert np.array_equal(o_h, o_d.cpu().numpy()) and np.array_equal(r_h, r_d.cpu().numpy())
        assert d_h.dtype == bool and np.array_equal(d_h, d_d.cpu().numpy().astype(bool))
    assert AircraftVecEnv.info_strings([0, 1, 2, 3, 4, 5]) == ["", "n", "c", "g", "w", "m"]
    # goal envs hand back dict observations
    gv = AircraftVecEnv("guidance-collision-avoidance-single-HER-v0", 64, n_intruders=3)
    d = gv.reset()
    assert set(d) == {"observation", "achieved_goal", "desired_goal"} and d["observation"].shape == (64, 18)
    for v in (venv, hv, venv2, gv):
        v.close()


def test_discrete9her_rewards_match_reference():
    """compute_reward (scalar, un-normalises its arguments in place) and compute_input_reward (stride-4 indexing quirk)
    of Simulators/SingleAircraftDiscrete9HEREnv.py:229-277 against values recorded from the reference."""
    import os
    from helpers import GOLDEN
    g = np.load(os.path.join(GOLDEN, "d9her_reward.npz"))
    env = _make("d9her", 8)
    assert env.observation_space.shape == (26,) and env.action_space.n == 9
    for i in range(len(g["r"])):
        assert env.compute_reward(g["ag"][i].copy(), g["g"][i].copy(), None) == g["r"][i]
        assert env.compute_input_reward(g["inputs"][i].copy()) == g["ri"][i]
    a0 = g["ag"][0].copy()
    env.compute_reward(a0, g["g"][0].copy(), None)
    assert np.array_equal(a0, g["ag0_after"])
    env.close()


def test_vec_monitor_equals_baselines_vec_monitor(tmp_path):
    """AircraftVecMonitor (device accumulators + episode ring, gca_monitor_update) against VecMonitor's arithmetic
    (vec_monitor.py:21-37: float32 eprets += rews, eplens += 1, record + restart on done) and the monitor.csv format."""
    import torch
    from gca_b200.vec_env import AircraftVecEnv
    from gca_b200.vec_monitor import AircraftVecMonitor
    B = 300
    venv = AircraftVecEnv("guidance-collision-avoidance-single-continuous-action-v0", B, n_intruders=20, seed=5)
    mon = AircraftVecMonitor(venv, filename=str(tmp_path / "run"))
    mon.reset()
    eprets, eplens = np.zeros(B, "f"), np.zeros(B, "i")
    want, got = [], []
    counts = {"steps": 0, "episodes": 0, "nmac": 0, "conflict_steps": 0, "goal": 0, "wall": 0, "maxsteps": 0}
    rng = np.random.RandomState(0)
    for t in range(400):
        a = torch.as_tensor(rng.uniform(-1, 1, (B, 2)).astype(np.float32), device="cuda")
        obs, rews, dones, infos = mon.step(a)
        rews, dones, codes = rews.cpu().numpy(), dones.cpu().numpy(), infos.cpu().numpy()
        counts["steps"] += B
        counts["episodes"] += int(dones.astype(bool).sum())
        for name, code in (("nmac", 1), ("conflict_steps", 2), ("goal", 3), ("wall", 4), ("maxsteps", 5)):
            counts[name] += int((codes == code).sum())
        eprets += rews
        eplens += 1
        for i in range(B):
            if dones[i]:
                want.append((i, float(eprets[i]), int(eplens[i])))
                eprets[i] = 0
                eplens[i] = 0
        if t % 97 == 0:
            got += mon.drain()
    got += mon.drain()
    assert len(want) > 50
    assert [(e["env"], e["r"], e["l"]) for e in got] == want
    assert all(e["t"] > 0 for e in got)
    assert mon.stats() == counts and counts["wall"] > 0 and counts["episodes"] == len(want)
    mon.close()
    lines = open(str(tmp_path / "run.monitor.csv")).read().splitlines()
    assert lines[0].startswith('# {"t_start": ') and lines[1] == "r,l,t" and len(lines) == 2 + len(want)
    r, l, t = lines[2].split(",")
    assert float(r) == want[0][1] and int(l) == want[0][2]


def test_vec_env_image_variant_with_frame_stack():
    """The registered stack id through the VecEnv interface: uint8 frames [B, 200, 200, k] on the device, k = 4 being
    baselines' VecFrameStack (roll, zero on done, newest frame last)."""
    import torch
    from gca_b200.vec_env import AircraftVecEnv
    B, k = 10, 4
    venv = AircraftVecEnv("guidance-collision-avoidance-single-stack-v0", B, n_intruders=12, seed=2, frame_stack=k)
    assert venv.observation_space.shape == (200, 200, k) and venv.observation_space.dtype == np.uint8
    obs = venv.reset()
    assert obs.shape == (B, 200, 200, k) and obs.dtype == torch.uint8 and obs.is_cuda
    stacked = obs.cpu().numpy().copy()
    assert not stacked[..., :-1].any() and (stacked[..., -1] == 255).mean() > 0.9
    rng = np.random.RandomState(0)
    for t in range(6):
        obs, rew, done, info = venv.step(torch.as_tensor(rng.randint(0, 9, B).astype(np.int32), device="cuda"))
        new = obs.cpu().numpy()
        d = done.cpu().numpy().astype(bool)
        stacked = np.roll(stacked, -1, axis=-1)
        stacked[d] = 0
        assert np.array_equal(new[..., :-1], stacked[..., :-1]), t
        stacked[..., -1] = new[..., -1]
    single = AircraftVecEnv("guidance-collision-avoidance-single-stack-v0", 3, n_intruders=5, seed=2)
    assert single.reset().shape == (3, 200, 200, 1)
    venv.close(); single.close()
