"""``basix.polynomials`` (python/basix/polynomials.py): the orthonormal polynomial sets are tabulated on the device."""

from enum import IntEnum

import numpy as np

from basix._basixcpp import NotBuiltNatively, PolysetType, polyset_dim  # noqa: F401
from basix._basixcpp import tabulate_polynomial_set_float32, tabulate_polynomial_set_float64


class PolynomialType(IntEnum):
    legendre = 0
    lagrange = 1
    bernstein = 2


def dim(ptype, celltype, degree):
    if int(ptype) != PolynomialType.legendre:
        raise NotBuiltNatively("only the orthonormal (Legendre) polynomial sets are built natively")
    return polyset_dim(celltype, PolysetType.standard, degree)


def tabulate_polynomials(ptype, celltype, degree, pts):
    """``basix.tabulate_polynomials``: (npolys, npts) values of the orthonormal set (polynomials.cpp: Legendre =
    polyset::tabulate with 0 derivatives)."""
    if int(ptype) != PolynomialType.legendre:
        raise NotBuiltNatively("only the orthonormal (Legendre) polynomial sets are built natively")
    pts = np.asarray(pts)
    fn = tabulate_polynomial_set_float32 if pts.dtype == np.float32 else tabulate_polynomial_set_float64
    return fn(celltype, PolysetType.standard, degree, 0, pts)[0]


def tabulate_polynomial_set(celltype, ptype, degree, nderiv, pts):
    """``basix.polynomials.tabulate_polynomial_set`` (python/basix/polynomials.py:207-231; polyset::tabulate,
    polyset.cpp:2914-2978): shape (nderivs, psize, npts), on the device."""
    pts = np.asarray(pts)
    if pts.dtype == np.float32:
        return tabulate_polynomial_set_float32(celltype, ptype, degree, nderiv, pts)
    if pts.dtype == np.float64:
        return tabulate_polynomial_set_float64(celltype, ptype, degree, nderiv, pts)
    raise RuntimeError(f"Unsupported data dtype: {pts.dtype}")
