"""Import-only stand-in for gym.envs.classic_control.rendering: Simulators/SingleAircraftMCTSRandIntruderEnv.py:13-28
derives a `Points` geom from `rendering.Geom` at import time.  Nothing is ever drawn here."""


class Geom(object):
    def __init__(self):
        self.attrs = []
