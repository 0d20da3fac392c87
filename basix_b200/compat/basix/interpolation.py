"""``basix.interpolation`` (python/basix/interpolation.py)."""

import basix_b200 as _bx


def compute_interpolation_operator(e0, e1):
    """``basix.compute_interpolation_operator(e0, e1)`` (interpolation.cpp:16-116), composed on the device."""
    return _bx.compute_interpolation_operator(getattr(e0, "_e", e0), getattr(e1, "_e", e1))
