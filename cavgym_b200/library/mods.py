"""Render-only hooks (reference library/mods.py:6-24).  Rendering (pyglet/OpenGL) is
out of the batched stepping path; the hooks stay importable and return plain data."""
import math


class FactoredLineStyle:
    def __init__(self, style, factor):
        self.style = style
        self.factor = factor


def make_circle(x, y, radius, res=30, filled=True):
    """Vertex list of a circle approximation (the reference wraps it in a gym rendering Geom)."""
    return [(x + math.cos(2 * math.pi * i / res) * radius, y + math.sin(2 * math.pi * i / res) * radius)
            for i in range(res)]
