"""Import path of Simulators/SingleAircraftRandomEnv.py (random ownship start)."""
from gca_b200.single import SingleAircraftRandomEnv  # noqa: F401
