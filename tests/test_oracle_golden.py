"""Pins the CPU oracle: the C restatement must reproduce, bit for bit, every trace recorded
from the unmodified reference (tests/golden/make_golden.py): reset, observation, reward, done,
info, conflict counter, the number of draws consumed and the full object state, step by step."""
import numpy as np
import pytest

from helpers import (FAST_TOL, GOAL_VARIANTS, GOLDEN_CASES, GOLDEN_N, GOLDEN_VARIANTS, assert_state_equal, fast_obs_tol, golden_actions,
                     golden_config, golden_state, load_trace)
from oracle import oracle as orc


def make_env(vk, n, g, auto_reset=False):
    cfg = golden_config(vk)
    B = g["tape"].shape[0]
    tape = np.nan_to_num(g["tape"], nan=0.0)
    return orc.OracleEnv(cfg, B, n, draws=0, trig=orc.TRIG_LIBM, tape=tape, auto_reset=auto_reset)


@pytest.mark.parametrize("vk,n", GOLDEN_CASES)
def test_reset_matches_reference(vk, n):
    g = load_trace(vk, n)
    env = make_env(vk, n, g)
    env.reset()
    assert np.array_equal(env.cursor, np.broadcast_to(g["cur_reset0"], env.cursor.shape))
    plain = np.nonzero(g["kind_id"] == 0)[0]          # traces whose start state is the untouched reset state
    want = golden_state(g, "s0_", plain)
    assert_state_equal(env.state, want, "reset %s n=%d" % (vk, n), rows=plain)
    assert np.array_equal(env.obs[plain], g["obs0"][plain])
    if vk in GOAL_VARIANTS:
        assert np.array_equal(env.achieved[plain], g["ag0"][plain])
        assert np.array_equal(env.desired[plain], g["dg0"][plain])


@pytest.mark.parametrize("vk,n", GOLDEN_CASES)
def test_free_running_replay_matches_reference(vk, n):
    g = load_trace(vk, n)
    env = make_env(vk, n, g)
    st = golden_state(g, "s0_")
    for k, v in st.items():
        env.state[k][...] = v
    env.cursor[...] = g["cur_reset0"]
    her = vk in GOAL_VARIANTS
    assert np.array_equal(env.observe(), g["obs0"])
    if her:
        assert np.array_equal(env.achieved, g["ag0"]) and np.array_equal(env.desired, g["dg0"])
    T = g["actions"].shape[1]
    sr_where = g["sr_where"]
    for t in range(T):
        what = "%s n=%d step %d" % (vk, n, t)
        assert np.array_equal(env.cursor, g["cur_before"][:, t]), what
        obs, rew, done, info = env.step(golden_actions(vk, g)[:, t])
        assert np.array_equal(info, g["event"][:, t]), what
        assert np.array_equal(done, g["done"][:, t]), what
        assert np.array_equal(rew, g["reward"][:, t]), what
        if vk == "d3her":                           # the 4th return value of Discrete3HER.step: dist_nearest_intruder
            assert np.array_equal(env.nearest, g["nearest"][:, t]), what
        assert np.array_equal(env.state["no_conflict"], g["no_conflict"][:, t]), what
        assert np.array_equal(env.cursor, g["cur_after"][:, t]), what
        assert np.array_equal(obs, g["obs"][:, t]), what
        if her:
            assert np.array_equal(env.achieved, g["ag"][:, t]) and np.array_equal(env.desired, g["dg"][:, t]), what
        want = golden_state(g, "sa_", (slice(None), t))
        assert_state_equal(env.state, want, what)
        if done.any():                                  # VecEnv-style reset of the finished envs
            env.reset(mask=done)
            rows = np.nonzero(done)[0]
            sel = [int(np.nonzero((sr_where[:, 0] == r) & (sr_where[:, 1] == t))[0][0]) for r in rows]
            want = golden_state(g, "sr_", sel)
            assert_state_equal(env.state, want, what + " reset", rows=rows)
            assert np.array_equal(env.obs[rows], g["reset_obs"][rows, t]), what
        assert np.array_equal(env.cursor, g["cur_after_reset"][:, t]), what


@pytest.mark.parametrize("vk,n", GOLDEN_CASES)
def test_fast_mode_replay_within_tolerance_of_reference(vk, n):
    """The fp32 ("fast") mode - the mode every headline number is measured on - against the reference traces, free
    running from the recorded start state with the recorded draws: flags, done, counters and draw counts identical,
    values inside helpers.FAST_TOL.  (The oracle's f32_positions variant is what the CUDA fast mode is bit-exact to,
    tests/test_gpu_parity.py::test_golden_replay_fast repeats this on the device.)"""
    g = load_trace(vk, n)
    B = g["tape"].shape[0]
    env = orc.OracleEnv(golden_config(vk), B, n, draws=0, trig=orc.TRIG_LIBM, tape=np.nan_to_num(g["tape"], nan=0.0),
                        f32_positions=True)
    for k, v in golden_state(g, "s0_").items():
        env.state[k][...] = v
    env.cursor[...] = g["cur_reset0"]
    f32 = lambda x: np.asarray(x).astype(np.float32).astype(np.float64)      # what the fast mode hands out / takes in
    acts = golden_actions(vk, g)
    if vk in ("env2", "her"):                   # continuous actions arrive as f32
        acts = f32(acts)
    otol = fast_obs_tol(vk)
    for t in range(acts.shape[1]):
        what = "%s n=%d step %d" % (vk, n, t)
        obs, rew, done, info = env.step(acts[:, t])
        assert np.array_equal(info, g["event"][:, t]) and np.array_equal(done, g["done"][:, t]), what
        assert np.array_equal(env.cursor, g["cur_after"][:, t]), what
        assert np.array_equal(env.state["no_conflict"], g["no_conflict"][:, t]), what
        want = golden_state(g, "sa_", (slice(None), t))
        assert np.array_equal(env.state["iflag"], want["iflag"]), what
        assert np.abs(env.state["own_pos"] - want["own_pos"]).max() <= FAST_TOL["pos"], what
        if n:
            assert np.abs(env.state["ipos"] - want["ipos"]).max() <= FAST_TOL["pos"], what
        if obs.shape[1]:
            assert np.abs(f32(obs) - g["obs"][:, t]).max() <= otol, what
        assert np.abs(f32(rew) - g["reward"][:, t]).max() <= FAST_TOL["reward"], what
        if vk in GOAL_VARIANTS:
            gtol = FAST_TOL["pos"] if vk == "dher" else FAST_TOL["obs"]     # DiscreteHER goals are raw pixels
            assert np.abs(f32(env.achieved) - g["ag"][:, t]).max() <= gtol and np.abs(f32(env.desired) - g["dg"][:, t]).max() <= gtol
        if vk == "d3her":
            assert np.abs(env.nearest - g["nearest"][:, t]).max() <= FAST_TOL["nearest"], what
        if done.any():
            env.reset(mask=done)
        assert np.array_equal(env.cursor, g["cur_after_reset"][:, t]), what


@pytest.mark.parametrize("vk,n", [("env", 80), ("env2", 3), ("her", 80), ("dher", 3), ("mcts", 80),
                                  ("mctsrnd", 80), ("mctsrnd", 3), ("stack", 80)])
def test_auto_reset_equals_step_then_reset(vk, n):
    """auto_reset folds the VecEnv contract (dummy_vec_env.py:52-55) into step."""
    g = load_trace(vk, n)
    env = make_env(vk, n, g, auto_reset=True)
    st = golden_state(g, "s0_")
    for k, v in st.items():
        env.state[k][...] = v
    env.cursor[...] = g["cur_reset0"]
    for t in range(g["actions"].shape[1]):
        obs, rew, done, info = env.step(golden_actions(vk, g)[:, t])
        assert np.array_equal(env.term_obs, g["obs"][:, t])
        assert np.array_equal(obs, g["reset_obs"][:, t])
        assert np.array_equal(done, g["done"][:, t]) and np.array_equal(rew, g["reward"][:, t])
        assert np.array_equal(env.cursor, g["cur_after_reset"][:, t])


def test_reference_kat_from_survey():
    """Config.intruder_size=3; np.random.seed(12345); actions [0,4,8,2,6] -
    replayed here from the same numpy stream (legacy RandomState is stable across versions)."""
    rs = np.random.RandomState(12345)
    tape = []
    for _ in range(3):                                  # x, y, speed, heading per spawn
        tape += [800 * rs.random_sample(), 800 * rs.random_sample()]
        tape += [rs.uniform(5 / 3, 8 / 3), rs.uniform(0, 2 * np.pi)]
    # no spawn was rejected for this seed (ownship at (50, 50)); then the goal, then 2 normals per step
    tape += list(rs.uniform(low=np.array([0, 0]), high=np.array([800, 800])))
    for _ in range(5):
        tape += [rs.normal(0, np.radians(2)), rs.normal(0, 2 / 30)]
    env = orc.OracleEnv(golden_config("env"), 1, 3, draws=0, trig=orc.TRIG_LIBM, tape=np.array([tape]))
    env.reset()
    rewards = [float(env.step(np.array([a]))[1][0]) for a in [0, 4, 8, 2, 6]]
    assert rewards == [-0.04669388650778382, -0.04695641387417, -0.04709211173559103, -0.04755891498538637,
                       -0.04781205946551119]
    assert env.state["own_pos"][0].tolist() == [np.float32(56.374012), np.float32(56.426086)]
    assert env.state["own_hs"][0].tolist() == [0.8657700806874236, 1.8536619964367387]
