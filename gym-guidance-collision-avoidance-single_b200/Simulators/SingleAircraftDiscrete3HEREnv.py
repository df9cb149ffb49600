"""Import path of the Discrete(3) training env (Algorithms/pytorch, Algorithms/A2C):
`from SingleAircraftDiscrete3HEREnv import SingleAircraftDiscrete3HEREnv` with Simulators/ on sys.path."""
from gca_b200.single import SingleAircraftDiscrete3HEREnv  # noqa: F401
