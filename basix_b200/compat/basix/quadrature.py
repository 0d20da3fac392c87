"""``basix.quadrature`` (python/basix/quadrature.py): Gauss-Jacobi rules from the library's native factory."""

import ctypes as C
from enum import IntEnum

import numpy as np

from basix._basixcpp import NotBuiltNatively, PolysetType


class QuadratureType(IntEnum):
    default = 0
    gauss_jacobi = 1
    gll = 2
    xiao_gimbutas = 3
    zienkiewicz_taylor = 20
    keast = 21
    strang_fix = 22


def make_quadrature(cell, degree, rule=QuadratureType.default, polyset_type=PolysetType.standard):
    """``basix.make_quadrature(cell, degree, rule, polyset_type)`` -> (points, weights).  The Gauss-Jacobi scheme
    (quadrature.cpp:215-372) is the one rule built natively; ``default`` resolves to it (the reference picks a tabulated
    Xiao-Gimbutas rule for low-degree simplices: a different point set of the same exactness)."""
    if int(rule) not in (QuadratureType.default, QuadratureType.gauss_jacobi) or int(polyset_type) != PolysetType.standard:
        raise NotBuiltNatively("make_quadrature: the Gauss-Jacobi scheme on the standard polyset is built natively")
    from basix_b200 import _capi
    from basix_b200.finite_element import _TDIM

    L = _capi.lib()
    n = C.c_size_t(0)
    _capi.check(L.bx_make_quadrature(int(cell), int(degree), None, None, C.byref(n)))
    pts = np.empty((n.value, _TDIM[int(cell)]), dtype=np.float64)
    wts = np.empty(n.value, dtype=np.float64)
    _capi.check(L.bx_make_quadrature(int(cell), int(degree), pts.ctypes.data, wts.ctypes.data, C.byref(n)))
    return pts, wts
