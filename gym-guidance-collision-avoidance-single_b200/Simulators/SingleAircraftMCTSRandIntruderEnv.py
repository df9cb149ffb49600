"""Import path of the env that Algorithms/MCTS/Agent_RandInt.py drives:
`from Simulators.SingleAircraftMCTSRandIntruderEnv import SingleAircraftEnv` (Agent_RandInt.py:9)."""
from gca_b200.single import SingleAircraftMCTSRandIntruderEnv as SingleAircraftEnv  # noqa: F401
