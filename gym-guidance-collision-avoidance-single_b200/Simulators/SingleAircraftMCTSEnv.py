"""Import path of the env that Algorithms/MCTS/Agent.py drives:
`from Simulators.SingleAircraftMCTSEnv import SingleAircraftEnv` (Agent.py:9)."""
from gca_b200.single import SingleAircraftMCTSEnv as SingleAircraftEnv  # noqa: F401
