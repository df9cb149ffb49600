"""Import path of the Simulators/ copy of the registered env (Config-driven reward row, info dict)."""
from gca_b200.single import SimSingleAircraftEnv as SingleAircraftEnv  # noqa: F401
