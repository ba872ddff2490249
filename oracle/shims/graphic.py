"""Shadows the reference's own pgtg/graphic.py (which needs matplotlib/PIL).
Only `create_map` is referenced (environment.py:750-769); rendering is out of scope."""


def create_map(*args, **kwargs):
    return None
