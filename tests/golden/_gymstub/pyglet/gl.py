"""Import-only stand-in for pyglet.gl."""
