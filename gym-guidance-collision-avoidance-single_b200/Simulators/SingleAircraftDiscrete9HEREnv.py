"""Import path of the env the repo's own learners train on (Algorithms/pytorch/dqn_her.py, Algorithms/A2C):
`from SingleAircraftDiscrete9HEREnv import SingleAircraftDiscrete9HEREnv` with Simulators/ on sys.path."""
from gca_b200.single import SingleAircraftDiscrete9HEREnv  # noqa: F401
