"""No-op stand-in for pyevtk.hl (reference use: dgfem/visualization.py:2)."""


def gridToVTK(*args, **kwargs):
    return None
