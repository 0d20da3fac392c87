"""``basix.lattice`` (python/basix/lattice.py): equispaced and GLL-warped point lattices."""

import ctypes as C
from enum import IntEnum

import numpy as np

from basix._basixcpp import NotBuiltNatively


class LatticeType(IntEnum):
    equispaced = 0
    gll = 1
    chebyshev = 2
    gl = 4
    chebyshev_plus_endpoints = 10
    gl_plus_endpoints = 11


class LatticeSimplexMethod(IntEnum):
    none = 0
    warp = 1
    isaac = 2
    centroid = 3


def create_lattice(celltype, n, ltype, exterior, method=LatticeSimplexMethod.none, dtype=np.float64):
    """``basix.create_lattice(celltype, n, ltype, exterior, method)`` -> (npts, tdim) array."""
    from basix_b200 import _capi

    npts = C.c_size_t(0)
    L = _capi.lib()
    rc = L.bx_create_lattice(int(celltype), int(n), int(ltype), int(bool(exterior)), int(method), None, C.byref(npts))
    if rc != 0:
        msg = L.bx_last_error().decode()
        raise (NotBuiltNatively if rc == 3 else RuntimeError)(msg)
    from basix_b200.finite_element import _TDIM

    out = np.empty((npts.value, _TDIM[int(celltype)]), dtype=np.float64)
    _capi.check(L.bx_create_lattice(int(celltype), int(n), int(ltype), int(bool(exterior)), int(method), out.ctypes.data,
                                    C.byref(npts)))
    return out.astype(dtype, copy=False)
