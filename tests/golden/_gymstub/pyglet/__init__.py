"""Import-only stand-in for pyglet (absent from this image): the reference's random-intruder MCTS env imports it at
module level for its render() only (Simulators/SingleAircraftMCTSRandIntruderEnv.py:11-12)."""
from . import gl  # noqa: F401
