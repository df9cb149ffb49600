"""Single-environment facade: the reference's gym classes served by a batch of one.

Each class keeps the reference's constructor (no required arguments), attributes, spaces,
`seed / reset / step / close / compute_reward` and return conventions (observation dtypes,
`int` vs `np.float64` rewards, the `info` string / dict) - see SURVEY.md 8(b) and the quirk list
Q11, Q13-Q18.  The state lives on the GPU (libgca, faithful mode: the reference's mixed f32/f64
arithmetic); there is no CPU implementation behind it.
"""
import math

import numpy as np

from . import abi, variants
from .batched import BatchedAircraftEnv, compute_reward as _device_compute_reward
from .spaces import Box, Dict, Discrete


class _SingleBase(object):
    VARIANT = None
    metadata = {"render.modes": []}
    reward_range = (-float("inf"), float("inf"))
    spec = None

    def _config_class(self):
        from gym_guidance_collision_avoidance_single.envs.config import Config
        return Config

    def __init__(self, device=0, seed=None, mode="faithful", time_limit=0, draws="philox", tape=None):
        self.Config = self._config_class()
        self.load_config()
        self.state = None
        self.viewer = None
        self._time_limit = int(time_limit)
        self._batch = BatchedAircraftEnv(self.VARIANT, 1, self.Config, n_intruders=self.intruder_size, mode=mode,
                                         draws=draws, device=device,
                                         seed=np.random.randint(2 ** 31) if seed is None else seed)
        if tape is not None:                          # parity tests: replay recorded numpy draws
            self._batch.set_tape(tape)
        if self._time_limit:
            self._batch.cfg.time_limit = self._time_limit
            self._batch.refresh_observation_params()
        self._faithful = mode == "faithful"
        self._build_spaces()
        self.position_range = Box(low=np.array([0, 0]), high=np.array([self.window_width, self.window_height]),
                                  dtype=np.float32)
        self.np_random = None
        self.seed(2)                                  # PKG/SingleAircraftEnv.py:43 (unused by the dynamics there)

    # PKG/SingleAircraftEnv.py:49-64
    def load_config(self):
        c = self.Config
        self.window_width = c.window_width
        self.window_height = c.window_height
        self.intruder_size = c.intruder_size
        self.EPISODES = c.EPISODES
        self.G = c.G
        self.tick = c.tick
        self.scale = c.scale
        self.minimum_separation = c.minimum_separation
        self.NMAC_dist = c.NMAC_dist
        self.horizon_dist = c.horizon_dist
        self.initial_min_dist = c.initial_min_dist
        self.goal_radius = c.goal_radius
        self.min_speed = c.min_speed
        self.max_speed = c.max_speed

    def _build_spaces(self):
        raise NotImplementedError

    # PKG/SingleAircraftEnv.py:45-47.  The returned list matches gym; the dynamics of the reference
    # draw from the global numpy stream whatever is passed here (Q1) - use reseed() to re-key ours.
    def seed(self, seed=None):
        self.np_random = np.random.RandomState(seed)
        return [seed]

    def reseed(self, seed):
        """Re-key the on-device Philox stream that drives this environment's randomness."""
        abi.check(self._batch.lib.gca_set_seed(self._batch._h, int(seed) & (2 ** 64 - 1)))

    @property
    def no_conflict(self):
        """Number of conflicts of the running episode (read by Algorithms/MCTS/Agent.py:52)."""
        return int(self._batch.get_state()["no_conflict"][0])

    @property
    def batch(self):
        return self._batch

    def _obs(self):
        b = self._batch
        return np.array(b.obs[0].cpu().numpy(), dtype=np.float64)

    def _format_obs(self):
        return self._obs()

    def reset(self):
        self._batch.refresh_observation_params()
        self._batch.reset()
        return self._format_obs()

    def _decode(self, action):
        return action

    def _info(self, code):
        return abi.INFO_STR[code]

    def step(self, action):
        import torch
        b = self._batch
        b.refresh_observation_params()
        a = self._decode(action)
        if b.continuous:
            t = torch.as_tensor(np.asarray(a, np.float64).reshape(1, 2), device=b.device).to(b.real)
        else:
            t = torch.as_tensor(np.asarray([int(a)], np.int32), device=b.device)
        b.step(t, auto_reset=False)
        code = int(b.info[0].item())
        done = bool(b.done[0].item())
        r = float(b.reward[0].item())
        if variants.reward_is_int(self.VARIANT, code):
            reward = int(r)                            # the reference returns Python ints for the constant rows (Q11)
        else:
            reward = np.float64(r)
        return self._format_obs(), reward, done, self._info(code)

    def render(self, mode="human"):
        raise NotImplementedError("rendering needs an OpenGL display in the reference; only SingleAircraftStackEnv "
                                  "produces frames here (device rasteriser)")

    def close(self):
        if getattr(self, "_batch", None) is not None:
            self._batch.close()
            self._batch = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class SingleAircraftEnv(_SingleBase):
    """Discrete-9 actions, vector observation (PKG/SingleAircraftEnv.py:14-184)."""
    VARIANT = "SingleAircraftEnv"

    def _build_spaces(self):
        dim = self.intruder_size * 4 + 8
        self.observation_space = Box(low=-1000, high=1000, shape=(dim,), dtype=np.float32)
        self.action_space = Discrete(9)


class SingleAircraft2Env(_SingleBase):
    """Continuous [-1,1]^2 actions (PKG/SingleAircraft2Env.py:12-176)."""
    VARIANT = "SingleAircraft2Env"

    def _build_spaces(self):
        dim = self.intruder_size * 4 + 8
        self.observation_space = Box(low=-1000, high=1000, shape=(dim,), dtype=np.float32)
        self.action_space = Box(low=-1, high=1, shape=(2,), dtype=float)

    def _decode(self, action):
        assert self.action_space.contains(action), "given action is in incorrect shape"   # :127
        return action


class _GoalBase(_SingleBase):
    def __init__(self, **kw):
        _SingleBase.__init__(self, **kw)

    def _format_obs(self):
        b = self._batch
        obs = np.array(b.obs[0].cpu().numpy(), dtype=np.float64)
        ag = b.achieved[0].cpu().numpy()
        dg = np.array(b.desired[0].cpu().numpy(), dtype=np.float64)
        # achieved_goal is an f32-valued quantity in the reference (Q13)
        return {"observation": obs, "achieved_goal": ag.astype(np.float32), "desired_goal": dg}

    def _build_spaces(self):
        # the reference resets inside the constructor to size its spaces (PKG/SingleAircraftHEREnv.py:32-39)
        obs = self.reset()
        self.observation_space = Dict(dict(
            desired_goal=Box(-np.inf, np.inf, shape=obs["achieved_goal"].shape, dtype="float32"),
            achieved_goal=Box(-np.inf, np.inf, shape=obs["achieved_goal"].shape, dtype="float32"),
            observation=Box(-np.inf, np.inf, shape=obs["observation"].shape, dtype="float32"),
        ))
        self._build_action_space()

    def compute_reward(self, achieved_goal, desired_goal, info):
        """Vectorised over leading axes like the reference; evaluated on the device."""
        import torch
        b = self._batch
        ag = torch.as_tensor(np.ascontiguousarray(achieved_goal), device=b.device)
        g = torch.as_tensor(np.ascontiguousarray(desired_goal), device=b.device)
        scalar = ag.dim() == 1
        if scalar:
            ag, g = ag[None], g[None]
        r = _device_compute_reward(ag, g, self.goal_radius, b.cfg.obs_kind).cpu().numpy()
        return r[0] if scalar else r


class SingleAircraftHEREnv(_GoalBase):
    """GoalEnv, continuous actions, dict observation (PKG/SingleAircraftHEREnv.py:12-196)."""
    VARIANT = "SingleAircraftHEREnv"

    def _build_action_space(self):
        self.action_space = Box(low=-1, high=1, shape=(2,), dtype=float)

    def _decode(self, action):
        if not self.action_space.contains(action):
            print("Warn: input action is", action)         # :142-143
        return action

    def _info(self, code):
        return {"result": abi.INFO_STR[code]}


class SingleAircraftDiscreteHEREnv(_GoalBase):
    """GoalEnv, Discrete(3) heading-only actions (PKG/SingleAircraftDiscreteHEREnv.py:12-186)."""
    VARIANT = "SingleAircraftDiscreteHEREnv"

    def _build_action_space(self):
        self.action_space = Discrete(3)

    def _format_obs(self):
        d = _GoalBase._format_obs(self)
        # raw pixel goals: drone.position.copy() is f32, goal.position.copy() is f64 (:131-132)
        return d

    def _info(self, code):
        return {}


class SingleAircraftMCTSEnv(_SingleBase):
    """The env Algorithms/MCTS/Agent.py drives (Simulators/SingleAircraftMCTSEnv.py): (a0, a1) tuple
    actions, raw un-normalised observation, Config-driven reward row."""
    VARIANT = "SingleAircraftMCTSEnv"

    def _config_class(self):
        from Simulators.config import Config
        return Config

    def _build_spaces(self):
        dim = self.intruder_size * 4 + 8
        self.observation_space = Box(low=-1000, high=1000, shape=(dim,), dtype=np.float32)
        self.action_space = Discrete(9)

    def _decode(self, action):
        if isinstance(action, (tuple, list, np.ndarray)):
            return int(action[0]) * 3 + int(action[1])
        return int(action)

    def step(self, action):
        import torch
        b = self._batch
        b.refresh_observation_params()
        t = torch.as_tensor(np.asarray([self._decode(action)], np.int32), device=b.device)
        b.step(t, auto_reset=False)
        code = int(b.info[0].item())
        r = float(b.reward[0].item())
        shaped = code == abi.INFO_NONE and not self.Config.sparse_reward
        reward = np.float64(r) if shaped else r            # Config values are Python floats (-10 / 10 ...)
        return self._obs(), reward, bool(b.done[0].item()), {"result": abi.INFO_STR[code]}


class SingleAircraftMCTSRandIntruderEnv(SingleAircraftMCTSEnv):
    """Simulators/SingleAircraftMCTSRandIntruderEnv.py, the env of Algorithms/MCTS/Agent_RandInt.py: the MCTS env whose
    intruders turn at random after every step (:166-174) and drift by Config.position_sigma (:183); the raw observation
    carries six entries per intruder (:133-140) although the declared space keeps 4 N + 8 (:50-51); `info` is the bare
    result string (:164)."""
    VARIANT = "SingleAircraftMCTSRandIntruderEnv"

    def load_config(self):
        super(SingleAircraftMCTSRandIntruderEnv, self).load_config()
        self.d_heading = self.Config.d_heading                  # :81-82
        self.position_sigma = self.Config.position_sigma

    def step(self, action):
        ob, reward, done, info = super(SingleAircraftMCTSRandIntruderEnv, self).step(action)
        return ob, reward, done, info["result"]


class SimSingleAircraftEnv(SingleAircraftMCTSEnv):
    """Simulators/SingleAircraftEnv.py: the registered env with the reward row of Simulators/config.py and an info dict
    (normalised vector observation, integer action 0..8)."""
    VARIANT = "SimSingleAircraftEnv"

    def _decode(self, action):
        return int(action)


class SingleAircraftRandomEnv(SimSingleAircraftEnv):
    """Simulators/SingleAircraftRandomEnv.py: as above with the ownship drawn at reset (:71-73)."""
    VARIANT = "SingleAircraftRandomEnv"


class SingleAircraftDiscrete9HEREnv(_GoalBase):
    """The training env of the repo's own learners (Simulators/SingleAircraftDiscrete9HEREnv.py, used by
    Algorithms/pytorch/dqn_her.py and Algorithms/A2C): random ownship start (:78-82), observation = ownship
    (x, y, vx, vy) + the Config.n nearest intruders (x, y, vx, vy, dist / Config.diagonal) (:106-144), dict goals,
    Discrete(9) actions, reward row of Simulators/config.py:37-43.  Needs Config.intruder_size > Config.n."""
    VARIANT = "SingleAircraftDiscrete9HEREnv"

    def _config_class(self):
        from Simulators.config import Config
        return Config

    def _build_spaces(self):
        obs = self.reset()                                  # the reference resets in its constructor (:31)
        # a flat Box of observation + desired goal, as the learners concatenate them (:42)
        self.observation_space = Box(low=-1000, high=1000, shape=(obs["observation"].shape[0] + 2,), dtype=np.float32)
        self.action_space = Discrete(9)

    def _info(self, code):
        return {"result": abi.INFO_STR[code]}

    def step(self, action):
        ob, reward, done, info = _SingleBase.step(self, action)
        code = abi.INFO_STR.index(info["result"])
        shaped = code == abi.INFO_NONE and not self.Config.sparse_reward
        return ob, (np.float64(reward) if shaped else float(reward)), done, info

    def unnormalize_position(self, position):               # :279-283 - in place, like the reference
        position[0] = position[0] * self.Config.window_width
        position[1] = position[1] * self.Config.window_height
        return position

    def compute_reward(self, achieved_goal, desired_goal, info):
        """:229-243 - scalar (one pair per call); un-normalises its arguments IN PLACE like the reference does."""
        c = self.Config
        achieved_goal = self.unnormalize_position(achieved_goal)
        desired_goal = self.unnormalize_position(desired_goal)
        d = np.linalg.norm((achieved_goal - desired_goal), axis=-1)
        if d < self.goal_radius:
            return c.goal_reward
        return c.step_penalty if c.sparse_reward else - d / 1200

    def compute_input_reward(self, new_inputs):
        """:245-277 - reward of a relabelled (observation + goal) vector.  The reference indexes the intruder entries
        with stride 4 (idx * 4 + 4) although each has 5 values: kept."""
        c = self.Config

        def metric(x1, y1, x2, y2):
            return math.sqrt((x1 - x2) ** 2 + (y1 - y2) ** 2)
        ownx = new_inputs[0] * c.window_width
        owny = new_inputs[1] * c.window_height
        gx = new_inputs[-2] * c.window_width
        gy = new_inputs[-1] * c.window_height
        dist_goal = metric(ownx, owny, gx, gy)
        if c.intruder_size != 0:
            for idx in range(c.n):
                intrux = new_inputs[idx * 4 + 4] * c.window_width
                intruy = new_inputs[idx * 4 + 5] * c.window_height
                dist_intruder = metric(ownx, owny, intrux, intruy)
                if dist_intruder < self.minimum_separation:
                    reward = c.conflict_penalty
                    if dist_intruder < self.NMAC_dist:
                        reward = c.NMAC_penalty
                    return reward
        if dist_goal < self.goal_radius:
            return c.goal_reward
        return c.step_penalty if c.sparse_reward else -dist_goal / 1200


class SingleAircraftDiscrete3HEREnv(SingleAircraftDiscrete9HEREnv):
    """Simulators/SingleAircraftDiscrete3HEREnv.py: the 9HER env with Discrete(3) heading-only actions (:407-411), the goal
    drawn 100 px inside the map (:349-353), the nearest-intruder term in the default reward (:225-232) - and step()
    returning dist_nearest_intruder where gym's info would be (:178)."""
    VARIANT = "SingleAircraftDiscrete3HEREnv"

    def _build_spaces(self):
        SingleAircraftDiscrete9HEREnv._build_spaces(self)
        self.action_space = Discrete(3)
        self.dist_nearest_intruder = 9999

    def step(self, action):
        ob, reward, done, _ = _SingleBase.step(self, action)
        self.dist_nearest_intruder = float(self._batch.nearest[0].item())
        return ob, np.float64(reward), done, self.dist_nearest_intruder


def _unused():  # keep math imported for parity with the reference module namespace
    return math.pi
