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


def gauss_jacobi_rule(alpha, npoints):
    """``basix.quadrature.gauss_jacobi_rule(alpha, npoints)`` (quadrature.cpp:4955-4965): points and weights on [0, 1] for
    the weight ``(1 - x)**alpha``."""
    from basix_b200 import _capi

    pts = np.empty(int(npoints), dtype=np.float64)
    wts = np.empty(int(npoints), dtype=np.float64)
    _capi.check(_capi.lib().bx_gauss_jacobi_rule(float(alpha), int(npoints), pts.ctypes.data, wts.ctypes.data))
    return pts, wts


def make_quadrature(cell, degree, rule=QuadratureType.default, polyset_type=PolysetType.standard):
    """``basix.make_quadrature(cell, degree, rule, polyset_type)`` -> (points, weights).  The Gauss-Jacobi scheme
    (quadrature.cpp:215-372) is the one rule built natively; ``default`` resolves to it (the reference picks a tabulated
    Xiao-Gimbutas rule for low-degree simplices: a different point set of the same exactness)."""
    if int(rule) not in (QuadratureType.default, QuadratureType.gauss_jacobi):
        raise NotBuiltNatively("make_quadrature: the Gauss-Jacobi scheme is the rule built natively")
    from basix_b200 import _capi
    from basix_b200.finite_element import _TDIM

    L = _capi.lib()
    n = C.c_size_t(0)
    _capi.check(L.bx_make_quadrature(int(cell), int(degree), None, None, C.byref(n)))
    pts = np.empty((n.value, _TDIM[int(cell)]), dtype=np.float64)
    wts = np.empty(n.value, dtype=np.float64)
    _capi.check(L.bx_make_quadrature(int(cell), int(degree), pts.ctypes.data, wts.ctypes.data, C.byref(n)))
    if int(polyset_type) == PolysetType.standard:
        return pts, wts
    # macroedge polysets are piecewise polynomials on the cell split at the edge midpoints: the rule is repeated on
    # every sub-cell (quadrature.cpp: make_macroedge_quadrature)
    c = int(cell)
    if c == 1:
        subs = [np.array([[0.0], [0.5]]), np.array([[0.5], [1.0]])]
    elif c == 2:
        a, b, d, e, f, g = (0, 0), (1, 0), (0, 1), (0.5, 0), (0.5, 0.5), (0, 0.5)
        subs = [np.array(v, dtype=float) for v in ((a, e, g), (e, b, f), (g, f, d), (e, f, g))]
    elif c in (4, 5):
        tdim = 2 if c == 4 else 3
        subs = []
        for corner in np.ndindex(*([2] * tdim)):
            o = 0.5 * np.array(corner, dtype=float)
            subs.append(np.vstack([o] + [o + 0.5 * np.eye(tdim)[k] for k in range(tdim)]))
    else:
        raise NotBuiltNatively("make_quadrature: macroedge rules are built for intervals, triangles, quadrilaterals and "
                               "hexahedra")
    out_p, out_w = [], []
    for v in subs:
        J = (v[1:] - v[0]).T  # affine map of the reference cell onto the sub-cell
        out_p.append(v[0] + pts @ J.T)
        out_w.append(wts * abs(np.linalg.det(J)))
    return np.vstack(out_p), np.concatenate(out_w)
