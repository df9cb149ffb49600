"""Same exports as the reference's envs/__init__.py:1-6 (classes are GPU-backed facades)."""
from gca_b200.single import (SingleAircraftEnv, SingleAircraft2Env, SingleAircraftHEREnv,  # noqa: F401
                             SingleAircraftDiscreteHEREnv)
from gym_guidance_collision_avoidance_single.envs.config import Config  # noqa: F401

try:
    from gca_b200.stack import SingleAircraftStackEnv  # noqa: F401
except ImportError:  # pragma: no cover
    pass
