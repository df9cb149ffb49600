"""Throw-away stand-in for the `gym` package (absent from this image).

TEST INFRASTRUCTURE ONLY: it exists so that tests/golden/make_golden.py can import the
UNMODIFIED reference from /root/reference in the build container and record golden
vectors.  Nothing in the product imports it.  Semantics restated from gym ~0.10-0.12
(the era of `timestep_limit`): Box.contains is shape-equal and low <= x <= high inclusive.
"""
from . import spaces  # noqa: F401
from . import utils   # noqa: F401
from . import envs    # noqa: F401


class Env(object):
    metadata = {}
    reward_range = (-float("inf"), float("inf"))
    action_space = None
    observation_space = None

    def step(self, action):
        raise NotImplementedError

    def reset(self):
        raise NotImplementedError

    def render(self, mode="human"):
        raise NotImplementedError

    def close(self):
        pass

    def seed(self, seed=None):
        return []


class GoalEnv(Env):
    def compute_reward(self, achieved_goal, desired_goal, info):
        raise NotImplementedError
