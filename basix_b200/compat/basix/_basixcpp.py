"""``basix._basixcpp``: the names of the reference's nanobind module (python/basix/wrappers/basix_wrappers.cpp:224-331,
basix_wrappers.h:467-475) over the C ABI of ``libbasix_b200.so``."""

import numpy as np

import basix_b200 as _bx
from basix_b200.finite_element import (CellType, DPCVariant, ElementFamily, LagrangeVariant, MapType,  # noqa: F401
                                       PolysetType, _TOPOLOGY)


class NotBuiltNatively(RuntimeError):
    """The requested object is outside what the B200 library builds (there is no fallback to another implementation)."""


# cell::geometry (cpp/basix/cell.cpp:20-50): vertices of the reference cells
_GEOMETRY = {
    0: [[]],
    1: [[0.0], [1.0]],
    2: [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0]],
    3: [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
    4: [[0.0, 0.0], [1.0, 0.0], [0.0, 1.0], [1.0, 1.0]],
    5: [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0],
        [0.0, 0.0, 1.0], [1.0, 0.0, 1.0], [0.0, 1.0, 1.0], [1.0, 1.0, 1.0]],
    6: [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [0.0, 0.0, 1.0], [1.0, 0.0, 1.0], [0.0, 1.0, 1.0]],
    7: [[0.0, 0.0, 0.0], [1.0, 0.0, 0.0], [0.0, 1.0, 0.0], [1.0, 1.0, 0.0], [0.0, 0.0, 1.0]],
}


def topology(celltype):
    return [[list(e) for e in row] for row in _TOPOLOGY[int(celltype)]]


def geometry(celltype):
    g = np.array(_GEOMETRY[int(celltype)], dtype=np.float64)
    return g.reshape(len(_GEOMETRY[int(celltype)]), -1)


def sub_entity_type(celltype, dim, index):
    c = int(celltype)
    tdim = len(_TOPOLOGY[c]) - 1
    if dim == 0:
        return CellType.point
    if dim == 1:
        return CellType.interval
    if dim == tdim:
        return CellType(c)
    n = len(_TOPOLOGY[c][2][index])
    return CellType.triangle if n == 3 else CellType.quadrilateral


def _entity_jacobians(celltype, e_dim):
    """cell::entity_jacobians (cell.cpp:662-689) through the library's geometry kernel (bit-identical)."""
    c = int(celltype)
    x = geometry(c)
    ents = _TOPOLOGY[c][e_dim]
    coords = np.stack([x[e[:e_dim + 1]] for e in ents])
    return _bx.geometry(coords, want=("J",))[0]


def cell_facet_jacobians(celltype):
    tdim = len(_TOPOLOGY[int(celltype)]) - 1
    if tdim not in (2, 3):
        raise RuntimeError("Facet jacobians not supported for this cell type.")
    return _entity_jacobians(celltype, tdim - 1)


def cell_edge_jacobians(celltype):
    if len(_TOPOLOGY[int(celltype)]) - 1 != 3:
        raise RuntimeError("Edge jacobians not supported for this cell type.")
    return _entity_jacobians(celltype, 1)


def index(p, q=None, r=None):
    return _bx.index(p, q, r)


def tabulate_polynomial_set_float64(celltype, ptype, d, n, x):
    """``_basixcpp.tabulate_polynomial_set_float64`` (basix_wrappers.h:467-475)."""
    return _bx.tabulate_polynomial_set(int(celltype), int(ptype), d, n, np.asarray(x, dtype=np.float64))


def tabulate_polynomial_set_float32(celltype, ptype, d, n, x):
    from basix_b200 import _capi

    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    P = np.empty((_bx.polyset_nderivs(celltype, n), _bx.polyset_dim(celltype, ptype, d), x.shape[0]), dtype=np.float32)
    _capi.check(_capi.lib().bx_polyset_tabulate_f32(int(celltype), int(ptype), d, n, x.ctypes.data, x.shape[0],
                                                    P.ctypes.data, _capi.BX_MEM_HOST, None))
    return P


def polyset_dim(celltype, ptype, d):
    return _bx.polyset_dim(celltype, ptype, d)


def create_element(family, celltype, degree, lvariant, dvariant, discontinuous, dof_ordering, dtype):
    """``_basixcpp.create_element(family, celltype, degree, lvariant, dvariant, discontinuous, dof_ordering, dtype_char)``
    (basix_wrappers.cpp:286-311).  One element object serves float32 and float64 arrays.  What the native factory does
    not build (status BX_ERR_UNSUPPORTED of bx_create_element) raises NotBuiltNatively."""
    import ctypes as C

    from basix_b200 import _capi
    _FE = FiniteElement_float64

    if np.dtype(dtype) not in (np.dtype(np.float32), np.dtype(np.float64)):
        raise TypeError("Unsupported finite element dtype.")
    order = np.ascontiguousarray(dof_ordering if dof_ordering is not None else [], dtype=np.int32)
    h = C.c_void_p()
    L = _capi.lib()
    rc = L.bx_create_element(int(family), int(celltype), int(degree), int(lvariant), int(dvariant), int(bool(discontinuous)),
                             order.ctypes.data if len(order) else None, len(order), C.byref(h))
    if rc != 0:
        msg = L.bx_last_error().decode()
        raise (NotBuiltNatively if rc == 3 else RuntimeError)(msg)
    lv = int(lvariant)
    if lv == 0 and int(family) == 1 and degree < 3:
        lv = int(LagrangeVariant.gll_warped)
    return _FE._from_handle(h, family=family, lagrange_variant=lv, dpc_variant=dvariant, discontinuous=discontinuous,
                            degree=degree)


def sub_entity_connectivity(celltype):
    """cell::sub_entity_connectivity (cell.cpp:143-330): for every sub-entity, the numbers of the entities of each
    dimension it is incident to (one contains the other), derived here from the topology."""
    topo = _TOPOLOGY[int(celltype)]
    out = []
    for ents in topo:
        row = []
        for verts in ents:
            v = set(verts)
            row.append([[i for i, w in enumerate(others) if set(w) <= v or v <= set(w)] for others in topo])
        out.append(row)
    return out


class FiniteElement_float64(_bx.FiniteElement):
    """The element object as the reference's nanobind class presents it (basix_wrappers.h:93-470): same calls as
    ``basix_b200.FiniteElement``; the sub-entity closure permutations act IN PLACE on the int32 array, as the raw binding
    does (the Python wrapper of ``basix.finite_element`` returns copies)."""

    def permute_subentity_closure(self, d, cell_or_entity_info, entity_type, entity_index=None):
        d2 = d.reshape(1, -1)
        self._closure(0, d2, np.asarray([int(cell_or_entity_info) & 0xFFFFFFFF], dtype=np.uint32), entity_type, entity_index)

    def permute_subentity_closure_inv(self, d, cell_or_entity_info, entity_type, entity_index=None):
        d2 = d.reshape(1, -1)
        self._closure(1, d2, np.asarray([int(cell_or_entity_info) & 0xFFFFFFFF], dtype=np.uint32), entity_type, entity_index)


FiniteElement_float32 = FiniteElement_float64
