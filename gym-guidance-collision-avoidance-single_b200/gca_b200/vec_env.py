"""Batched, device-resident VecEnv with the OpenAI-baselines interface.

Replaces `DummyVecEnv / SubprocVecEnv / ShmemVecEnv` over B Python env objects
(baselines/common/vec_env/__init__.py:26-135, dummy_vec_env.py:46-63) by one kernel launch per
step: `reset()`, `step_async(actions)`, `step_wait() -> (obs, rews, dones, infos)`, `step()`,
`close()`, `num_envs / observation_space / action_space`, auto-reset on done (the returned
observation of a finished env is reset()'s, Q20) and the AlreadyStepping / NotStepping errors.

By default everything stays on the GPU (torch tensors, no host sync per step); `host=True`
returns numpy arrays through pinned buffers (the end-to-end path measured by bench.py).
"""
import numpy as np

from . import abi
from .batched import BatchedAircraftEnv
from .spaces import Box, Dict, Discrete


class AlreadySteppingError(Exception):
    def __init__(self):
        Exception.__init__(self, "already running an async step")


class NotSteppingError(Exception):
    def __init__(self):
        Exception.__init__(self, "not running an async step")


_CLASS_OF = {
    "guidance-collision-avoidance-single-v0": "SingleAircraftEnv",
    "guidance-collision-avoidance-single-continuous-action-v0": "SingleAircraft2Env",
    "guidance-collision-avoidance-single-stack-v0": "SingleAircraftStackEnv",
    "guidance-collision-avoidance-single-HER-v0": "SingleAircraftHEREnv",
    "guidance-collision-avoidance-single-Discrete-HER-v0": "SingleAircraftDiscreteHEREnv",
}


class AircraftVecEnv(object):
    def __init__(self, env, num_envs, config=None, n_intruders=None, mode="fast", device=0, seed=0, env_id0=0,
                 host=False, time_limit=None, frame_stack=1, sprites=None, sprite_dir=None):
        """frame_stack / sprites / sprite_dir apply to the image variant only (SingleAircraftStackEnv): k > 1 is baselines'
        VecFrameStack(venv, k) (common/vec_env/vec_frame_stack.py) with the frames kept in a ring on the device;
        sprite_dir points at the reference's images (gca_b200/sprites.py)."""
        variant = _CLASS_OF.get(env, env)
        registered = env in _CLASS_OF
        self._image = None
        if variant == "SingleAircraftStackEnv":
            self._init_image(num_envs, config, n_intruders, mode, device, seed, env_id0, host, frame_stack, sprites, sprite_dir)
            return
        if config is None:
            from .variants import default_config_class
            config = default_config_class(variant)
        self.variant = variant
        self.num_envs = int(num_envs)
        self.host = bool(host)
        self.batch = BatchedAircraftEnv(variant, num_envs, config, n_intruders=n_intruders, mode=mode, device=device,
                                        seed=seed, env_id0=env_id0)
        if time_limit is None:
            time_limit = 10000 if registered else 0       # timestep_limit of the registered ids (Q19)
        if time_limit:
            self.batch.cfg.time_limit = int(time_limit)
            self.batch.refresh_observation_params()
        n = self.batch.n_intruders
        b = self.batch
        if b.is_goal_env:
            self.observation_space = Dict(dict(
                desired_goal=Box(-np.inf, np.inf, shape=(2,), dtype="float32"),
                achieved_goal=Box(-np.inf, np.inf, shape=(2,), dtype="float32"),
                observation=Box(-np.inf, np.inf, shape=(b.obs_dim,), dtype="float32")))
        else:   # (the random-intruder env declares 4 N + 8 and returns 6 N + 8 values, Simulators/SingleAircraftMCTSRandIntruderEnv.py:52: kept)
            self.observation_space = Box(low=-1000, high=1000, shape=(4 * n + 8,), dtype=np.float32)
        if b.continuous:
            self.action_space = Box(low=-1, high=1, shape=(2,), dtype=np.float32)
        else:
            # Discrete(3): PKG/SingleAircraftDiscreteHEREnv.py and Simulators/SingleAircraftDiscrete3HEREnv.py (heading only)
            self.action_space = Discrete(3 if b.cfg.action_kind in (abi.ACT_DISCRETE3, abi.ACT_DISCRETE3_HEADING) else 9)
        self._pending = None
        self.closed = False

    def _init_image(self, num_envs, config, n_intruders, mode, device, seed, env_id0, host, frame_stack, sprites, sprite_dir=None):
        from .stack import ImageBatch
        if host:
            raise ValueError("the image variant keeps its frames on the device (host=False)")
        self.variant = "SingleAircraftStackEnv"
        self.num_envs, self.host = int(num_envs), False
        self._image = ImageBatch(num_envs, config, n_intruders=n_intruders, frame_stack=frame_stack, device=device,
                                 seed=seed, env_id0=env_id0, mode=mode, sprites=sprites, sprite_dir=sprite_dir)
        self.batch = self._image.batch
        k = self._image.k
        # PKG/SingleAircraftStackEnv.py:34-35 (200 x 200 x 1 uint8), repeated k times on the last axis by VecFrameStack
        self.observation_space = Box(low=0, high=255, shape=(self._image.H, self._image.W, k), dtype=np.uint8)
        self.action_space = Discrete(9)
        self._pending = None
        self.closed = False

    def _image_obs(self):
        return self._image.stacked() if self._image.k > 1 else self._image.frame()

    # ------------------------------------------------------------------ VecEnv interface
    def _pack_obs(self, obs):
        b = self.batch
        if not b.is_goal_env:
            return obs
        if self.host:
            views = b.last_host_views                    # the pinned buffer set the last host call filled
            return {"observation": views["obs"], "achieved_goal": views["achieved"], "desired_goal": views["desired"]}
        return {"observation": b.obs, "achieved_goal": b.achieved, "desired_goal": b.desired}

    def reset(self):
        if self._image is not None:
            self._image.reset()
            return self._image_obs()
        if self.host:
            return self._pack_obs(self.batch.reset_host())
        return self._pack_obs(self.batch.reset())

    def step_async(self, actions):
        if self._pending is not None:
            raise AlreadySteppingError()
        if self._image is not None:
            _, rew, done, info = self._image.step(actions, auto_reset=True)
            self._pending = ("img", (rew, done, info))
        elif self.host:
            # gca_step_host_begin: upload, step and download are enqueued; step_wait blocks on the download
            self.batch.step_host_begin(np.asarray(actions), auto_reset=True)
            self._pending = ("host", None)
        else:
            # the launch is asynchronous on the current CUDA stream: this IS the async half
            self._pending = ("dev", self.batch.step(actions, auto_reset=True))

    def step_wait(self):
        if self._pending is None:
            raise NotSteppingError()
        kind, payload = self._pending
        self._pending = None
        if kind == "img":
            rew, done, info = payload
            return self._image_obs(), rew, done, info
        if kind == "host":
            obs, rew, done, info = self.batch.step_host_wait()
            return self._pack_obs(obs), rew, done.astype(bool), info
        obs, rew, done, info = payload
        return self._pack_obs(obs), rew, done, info

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        if not self.closed:
            self.batch.close()
            self.closed = True

    def get_images(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    @property
    def unwrapped(self):
        return self

    @staticmethod
    def info_strings(info_codes):
        """Decode the uint8 result codes into the reference's info strings ('' n c g w m)."""
        return [abi.INFO_STR[int(c)] for c in np.asarray(info_codes).ravel()]
