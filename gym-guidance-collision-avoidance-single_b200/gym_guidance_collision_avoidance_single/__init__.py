"""Drop-in package path of the reference: importing it registers the five environment ids
(gym_guidance_collision_avoidance_single/__init__.py:6-44 of the reference) with `gym` /
`gymnasium` when one of them is installed, and always with the built-in registry below
(`make(id)` applies the reference's timestep_limit=10000 like gym.make would, Q19).
"""
import logging

logger = logging.getLogger(__name__)

_SPECS = [
    ("guidance-collision-avoidance-single-v0", "SingleAircraftEnv"),
    ("guidance-collision-avoidance-single-continuous-action-v0", "SingleAircraft2Env"),
    ("guidance-collision-avoidance-single-stack-v0", "SingleAircraftStackEnv"),
    ("guidance-collision-avoidance-single-HER-v0", "SingleAircraftHEREnv"),
    ("guidance-collision-avoidance-single-Discrete-HER-v0", "SingleAircraftDiscreteHEREnv"),
]

registry = {}
for _id, _cls in _SPECS:
    registry[_id] = dict(entry_point="gym_guidance_collision_avoidance_single.envs:%s" % _cls, timestep_limit=10000,
                         reward_threshold=10.0, nondeterministic=False)


def _register_with(module):
    for _id, spec in registry.items():
        try:
            try:
                module.register(id=_id, entry_point=spec["entry_point"], max_episode_steps=spec["timestep_limit"],
                                reward_threshold=spec["reward_threshold"], nondeterministic=False)
            except TypeError:
                module.register(id=_id, entry_point=spec["entry_point"], timestep_limit=spec["timestep_limit"],
                                reward_threshold=spec["reward_threshold"], nondeterministic=False)
        except Exception as e:  # already registered, incompatible version ...
            logger.debug("gym registration of %s skipped: %s", _id, e)


for _name in ("gym.envs.registration", "gymnasium.envs.registration"):
    try:  # pragma: no cover - neither is installed in the build image
        import importlib
        _register_with(importlib.import_module(_name))
    except ImportError:
        pass


def make(env_id, **kwargs):
    """gym.make for the registered ids without gym: constructs the class with the TimeLimit of the spec."""
    import importlib
    spec = registry[env_id]
    mod, cls = spec["entry_point"].split(":")
    kwargs.setdefault("time_limit", spec["timestep_limit"])
    return getattr(importlib.import_module(mod), cls)(**kwargs)
