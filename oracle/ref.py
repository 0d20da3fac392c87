"""TEST INFRASTRUCTURE ONLY -- never part of the product path.

ctypes driver for ``oracle/_ref/libbasix_ref.so``: the unmodified reference
C++ core (FEniCS/basix) compiled by ``oracle/build_ref.sh`` plus the
``extern "C"`` shim in ``oracle/ref_shim.cpp``.  Only ``tests/``,
``__graft_entry__.smoke()`` and ``bench.py``'s cpu_baseline / ``--impl
reference`` legs may import this module.

The class mirrors the reference's Python surface
(python/basix/finite_element.py:50-592) closely enough for parity tests to
read like the reference's own tests.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_ref", "libbasix_ref.so")

_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def lib():
    global _lib
    if _lib is None:
        if not available():
            raise RuntimeError(f"reference oracle not built: {LIB_PATH} (run oracle/build_ref.sh)")
        L = C.CDLL(LIB_PATH)
        L.bxref_last_error.restype = C.c_char_p
        L.bxref_create_element.restype = C.c_void_p
        L.bxref_create_element.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_int]
        L.bxref_destroy_element.argtypes = [C.c_void_p]
        L.bxref_element_info.argtypes = [C.c_void_p, C.c_void_p]
        L.bxref_value_shape.argtypes = [C.c_void_p, C.c_void_p]
        L.bxref_coeffs.argtypes = [C.c_void_p, C.c_void_p]
        L.bxref_dof_ordering.argtypes = [C.c_void_p, C.c_void_p]
        L.bxref_num_entities.argtypes = [C.c_void_p, C.c_int]
        L.bxref_entity_dofs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.bxref_sub_entity_type.argtypes = [C.c_int] * 3
        L.bxref_entity_transformations.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.bxref_polyset_dim.argtypes = [C.c_int] * 3
        L.bxref_polyset_nderivs.argtypes = [C.c_int] * 2
        L.bxref_polyset_tabulate.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.bxref_tabulate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p,
                                     C.c_size_t, C.c_int]
        L.bxref_map.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_size_t] * 5 + [
            C.c_void_p, C.c_void_p, C.c_int]
        L.bxref_T_apply_f64.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                        C.c_void_p, C.c_int]
        L.bxref_T_apply_f32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                        C.c_void_p]
        L.bxref_T_apply_i32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int,
                                        C.c_void_p]
        L.bxref_permute_i32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p,
                                        C.c_int]
        L.bxref_base_transformations.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bxref_entity_closure_dofs.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.bxref_points.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bxref_interpolation_matrix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.bxref_permute_subentity_closure_i32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t,
                                                          C.c_void_p, C.c_int, C.c_int]
        L.bxref_entity_jacobians.argtypes = [C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.bxref_sub_entity_geometry.argtypes = [C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.bxref_compute_interpolation_operator.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
        L.bxref32_create_element.restype = C.c_void_p
        L.bxref32_create_element.argtypes = [C.c_int] * 6 + [C.c_void_p, C.c_int]
        L.bxref32_destroy_element.argtypes = [C.c_void_p]
        L.bxref32_polyset_tabulate.argtypes = [C.c_int] * 4 + [C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.bxref32_tabulate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_void_p]
        L.bxref32_map.argtypes = [C.c_void_p, C.c_int] + [C.c_void_p] * 4 + [C.c_size_t] * 5 + [C.c_void_p, C.c_void_p]
        L.bxref32_T_apply.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p]
        _lib = L
    return _lib


def _check(rc):
    if rc != 0:
        raise RuntimeError(lib().bxref_last_error().decode())


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


T_OPS = {"T_apply": 0, "Tt_apply": 1, "Tt_inv_apply": 2, "Tinv_apply": 3, "Tt_apply_right": 4,
         "Tinv_apply_right": 5, "T_apply_right": 6, "Tt_inv_apply_right": 7}


def polyset_dim(cell: int, ptype: int, d: int) -> int:
    return lib().bxref_polyset_dim(int(cell), int(ptype), int(d))


def polyset_nderivs(cell: int, n: int) -> int:
    return lib().bxref_polyset_nderivs(int(cell), int(n))


def tabulate_polynomial_set(cell: int, ptype: int, d: int, n: int, x: np.ndarray) -> np.ndarray:
    """reference polyset::tabulate -> P[nderivs][psize][npts] (polyset.cpp:2914-2992)."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    npts, tdim = x.shape
    P = np.empty((polyset_nderivs(cell, n), polyset_dim(cell, ptype, d), npts), dtype=np.float64)
    _check(lib().bxref_polyset_tabulate(int(cell), int(ptype), int(d), int(n), _ptr(x), npts, tdim, _ptr(P)))
    return P


def entity_jacobians(cell: int, e_dim: int) -> np.ndarray:
    """reference cell::entity_jacobians (cell.cpp:662-689): shape (n_entities, tdim, e_dim)."""
    shp = np.zeros(3, dtype=np.int32)
    _check(lib().bxref_entity_jacobians(int(cell), int(e_dim), _ptr(shp), None))
    out = np.empty(tuple(int(v) for v in shp), dtype=np.float64)
    _check(lib().bxref_entity_jacobians(int(cell), int(e_dim), _ptr(shp), _ptr(out)))
    return out


def sub_entity_geometry(cell: int, dim: int, index: int) -> np.ndarray:
    """reference cell::sub_entity_geometry: vertex coordinates of one sub-entity, shape (nvertices, tdim)."""
    shp = np.zeros(2, dtype=np.int32)
    _check(lib().bxref_sub_entity_geometry(int(cell), int(dim), int(index), _ptr(shp), None))
    out = np.empty(tuple(int(v) for v in shp), dtype=np.float64)
    _check(lib().bxref_sub_entity_geometry(int(cell), int(dim), int(index), _ptr(shp), _ptr(out)))
    return out


def compute_interpolation_operator(e0: "RefElement", e1: "RefElement") -> np.ndarray:
    """reference basix::compute_interpolation_operator (interpolation.cpp:16-116)."""
    shp = np.zeros(2, dtype=np.int32)
    _check(lib().bxref_compute_interpolation_operator(e0._h, e1._h, _ptr(shp), None))
    out = np.empty((int(shp[0]), int(shp[1])), dtype=np.float64)
    _check(lib().bxref_compute_interpolation_operator(e0._h, e1._h, _ptr(shp), _ptr(out)))
    return out


class RefElement:
    """Reference ``basix::FiniteElement<double>`` handle."""

    def __init__(self, family, cell, degree, lagrange_variant=0, dpc_variant=0, discontinuous=False,
                 dof_ordering=None):
        L = lib()
        ord_ = np.asarray(dof_ordering if dof_ordering is not None else [], dtype=np.int32)
        self._h = L.bxref_create_element(int(family), int(cell), int(degree), int(lagrange_variant),
                                         int(dpc_variant), int(bool(discontinuous)), _ptr(ord_), len(ord_))
        if not self._h:
            raise RuntimeError(L.bxref_last_error().decode())
        info = np.zeros(12, dtype=np.int32)
        L.bxref_element_info(self._h, _ptr(info))
        (self.cell_type, self.polyset_type, self.tdim, self.dim, self.value_size, self.map_type,
         self.embedded_superdegree, perms, ident, self.coeff_cols, self.degree, n_ord) = [int(v) for v in info]
        self.family = int(family)
        self.lagrange_variant = int(lagrange_variant)
        self.dpc_variant = int(dpc_variant)
        self.discontinuous = bool(discontinuous)
        self.dof_transformations_are_permutations = bool(perms)
        self.dof_transformations_are_identity = bool(ident)
        shp = np.zeros(8, dtype=np.int32)
        k = L.bxref_value_shape(self._h, _ptr(shp))
        self.value_shape = tuple(int(s) for s in shp[:k])
        self.coefficient_matrix = np.empty((self.dim, self.coeff_cols), dtype=np.float64)
        L.bxref_coeffs(self._h, _ptr(self.coefficient_matrix))
        self.dof_ordering = np.zeros(n_ord, dtype=np.int32)
        if n_ord:
            L.bxref_dof_ordering(self._h, _ptr(self.dof_ordering))
        self.entity_dofs = []
        for d in range(self.tdim + 1):
            row = []
            for i in range(L.bxref_num_entities(self._h, d)):
                n = L.bxref_entity_dofs(self._h, d, i, None)
                buf = np.zeros(max(n, 1), dtype=np.int32)
                L.bxref_entity_dofs(self._h, d, i, _ptr(buf))
                row.append([int(v) for v in buf[:n]])
            self.entity_dofs.append(row)
        self.subentity_types = [
            [L.bxref_sub_entity_type(self.cell_type, d, i) for i in range(len(self.entity_dofs[d]))]
            for d in range(self.tdim + 1)]
        shp = np.zeros(2, dtype=np.int32)
        L.bxref_points(self._h, _ptr(shp), None)
        self.points = np.empty((int(shp[0]), int(shp[1])), dtype=np.float64)
        L.bxref_points(self._h, _ptr(shp), _ptr(self.points))
        L.bxref_interpolation_matrix(self._h, _ptr(shp), None)
        self.interpolation_matrix = np.empty((int(shp[0]), int(shp[1])), dtype=np.float64)
        L.bxref_interpolation_matrix(self._h, _ptr(shp), _ptr(self.interpolation_matrix))
        self.entity_transformations = {}
        for ct in range(8):
            shape = np.zeros(3, dtype=np.int32)
            if L.bxref_entity_transformations(self._h, ct, _ptr(shape), None):
                out = np.empty(tuple(int(s) for s in shape), dtype=np.float64)
                L.bxref_entity_transformations(self._h, ct, _ptr(shape), _ptr(out))
                self.entity_transformations[ct] = out

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().bxref_destroy_element(self._h)
                self._h = None
        except Exception:
            pass

    # --- hot path -----------------------------------------------------------------------------
    def tabulate(self, nd: int, x: np.ndarray, chunk: int = 0, nthreads: int = 1) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float64)
        if x.ndim != 2:
            raise ValueError("x must be 2-D")
        npts = x.shape[0]
        nder = polyset_nderivs(self.cell_type, nd)
        out = np.empty((nder, npts, self.dim, self.value_size), dtype=np.float64)
        _check(lib().bxref_tabulate(self._h, int(nd), _ptr(x), npts, x.shape[1], _ptr(out), int(chunk),
                                    int(nthreads)))
        return out

    def _map(self, pull, U, J, detJ, K, nthreads=1):
        U = np.ascontiguousarray(U, dtype=np.float64)
        J = np.ascontiguousarray(J, dtype=np.float64)
        detJ = np.ascontiguousarray(detJ, dtype=np.float64)
        K = np.ascontiguousarray(K, dtype=np.float64)
        nc, rows, in_vs = U.shape
        gdim, tdim = J.shape[1], J.shape[2]
        ovs = C.c_int(0)
        L = lib()
        if nc == 0:
            return np.empty((0, rows, 0))
        # dry run on one cell for the output value size
        _check(L.bxref_map(self._h, pull, _ptr(U), _ptr(J), _ptr(detJ), _ptr(K), 1, rows, in_vs, gdim, tdim,
                           None, C.byref(ovs), 1))
        out = np.empty((nc, rows, ovs.value), dtype=np.float64)
        _check(L.bxref_map(self._h, pull, _ptr(U), _ptr(J), _ptr(detJ), _ptr(K), nc, rows, in_vs, gdim, tdim,
                           _ptr(out), C.byref(ovs), int(nthreads)))
        return out

    def push_forward(self, U, J, detJ, K, nthreads=1):
        return self._map(0, U, J, detJ, K, nthreads)

    def pull_back(self, u, J, detJ, K, nthreads=1):
        return self._map(1, u, J, detJ, K, nthreads)

    def T_apply_batch(self, op: str, data: np.ndarray, n: int, cell_info: np.ndarray, nthreads: int = 1):
        """In place; data[nc][cell_size]; one reference call per cell."""
        assert data.flags.c_contiguous and data.ndim == 2
        ci = np.ascontiguousarray(cell_info, dtype=np.uint32)
        nc, cs = data.shape
        assert ci.shape == (nc,)
        L = lib()
        if data.dtype == np.float64:
            _check(L.bxref_T_apply_f64(self._h, T_OPS[op], _ptr(data), nc, cs, int(n), _ptr(ci), int(nthreads)))
        elif data.dtype == np.float32:
            _check(L.bxref_T_apply_f32(self._h, T_OPS[op], _ptr(data), nc, cs, int(n), _ptr(ci)))
        elif data.dtype == np.int32:
            _check(L.bxref_T_apply_i32(self._h, T_OPS[op], _ptr(data), nc, cs, int(n), _ptr(ci)))
        else:
            raise TypeError(data.dtype)
        return data

    def T_apply(self, data: np.ndarray, n: int, cell_info: int):
        flat = data.reshape(1, -1)
        self.T_apply_batch("T_apply", flat, n, np.array([cell_info], dtype=np.uint32))
        return data

    def permute_batch(self, d: np.ndarray, cell_info: np.ndarray, inverse: bool = False, nthreads: int = 1):
        assert d.dtype == np.int32 and d.flags.c_contiguous and d.ndim == 2
        ci = np.ascontiguousarray(cell_info, dtype=np.uint32)
        _check(lib().bxref_permute_i32(self._h, int(inverse), _ptr(d), d.shape[0], d.shape[1], _ptr(ci),
                                       int(nthreads)))
        return d

    def base_transformations(self) -> np.ndarray:
        shp = np.zeros(3, dtype=np.int32)
        lib().bxref_base_transformations(self._h, _ptr(shp), None)
        out = np.empty(tuple(int(v) for v in shp), dtype=np.float64)
        lib().bxref_base_transformations(self._h, _ptr(shp), _ptr(out))
        return out

    @property
    def entity_closure_dofs(self):
        L, out = lib(), []
        for d in range(self.tdim + 1):
            row = []
            for i in range(len(self.entity_dofs[d])):
                n = L.bxref_entity_closure_dofs(self._h, d, i, None)
                buf = np.zeros(max(n, 1), dtype=np.int32)
                L.bxref_entity_closure_dofs(self._h, d, i, _ptr(buf))
                row.append([int(v) for v in buf[:n]])
            out.append(row)
        return out

    def permute_subentity_closure_batch(self, d: np.ndarray, info: np.ndarray, entity_type: int,
                                        entity_index: int | None = None, inverse: bool = False):
        """FiniteElement::permute_subentity_closure{,_inv} per row of d (finite-element.h:860-1035); ``info`` holds
        entity_info when ``entity_index`` is None, else cell_info."""
        assert d.dtype == np.int32 and d.flags.c_contiguous and d.ndim == 2
        ci = np.ascontiguousarray(info, dtype=np.uint32)
        _check(lib().bxref_permute_subentity_closure_i32(self._h, int(inverse), _ptr(d), d.shape[0], d.shape[1], _ptr(ci),
                                                         int(entity_type), -1 if entity_index is None else int(entity_index)))
        return d


# ---- the reference's float32 instantiation (FiniteElement<float>, finite-element.cpp:2051) ------------------------
def tabulate_polynomial_set_f32(cell: int, ptype: int, d: int, n: int, x: np.ndarray) -> np.ndarray:
    """reference polyset::tabulate<float> (polyset.cpp:2995-2997)."""
    x = np.ascontiguousarray(x, dtype=np.float32)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    npts, tdim = x.shape
    P = np.empty((polyset_nderivs(cell, n), polyset_dim(cell, ptype, d), npts), dtype=np.float32)
    _check(lib().bxref32_polyset_tabulate(int(cell), int(ptype), int(d), int(n), _ptr(x), npts, tdim, _ptr(P)))
    return P


class RefElement32:
    """Reference ``basix::FiniteElement<float>`` handle: the hot-path calls on float32 arrays."""

    def __init__(self, family, cell, degree, lagrange_variant=0, dpc_variant=0, discontinuous=False, dof_ordering=None):
        L = lib()
        ord_ = np.asarray(dof_ordering if dof_ordering is not None else [], dtype=np.int32)
        self._h = L.bxref32_create_element(int(family), int(cell), int(degree), int(lagrange_variant), int(dpc_variant),
                                           int(bool(discontinuous)), _ptr(ord_), len(ord_))
        if not self._h:
            raise RuntimeError(L.bxref_last_error().decode())
        d = RefElement(family, cell, degree, lagrange_variant, dpc_variant, discontinuous, dof_ordering)
        self.cell_type, self.dim, self.value_size = d.cell_type, d.dim, d.value_size

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                lib().bxref32_destroy_element(self._h)
                self._h = None
        except Exception:
            pass

    def tabulate(self, nd: int, x: np.ndarray) -> np.ndarray:
        x = np.ascontiguousarray(x, dtype=np.float32)
        out = np.empty((polyset_nderivs(self.cell_type, nd), x.shape[0], self.dim, self.value_size), dtype=np.float32)
        _check(lib().bxref32_tabulate(self._h, int(nd), _ptr(x), x.shape[0], x.shape[1], _ptr(out)))
        return out

    def _map(self, pull, U, J, detJ, K):
        U, J, detJ, K = (np.ascontiguousarray(a, dtype=np.float32) for a in (U, J, detJ, K))
        nc, rows, in_vs = U.shape
        gdim, tdim = J.shape[1], J.shape[2]
        ovs = C.c_int(0)
        L = lib()
        _check(L.bxref32_map(self._h, pull, _ptr(U), _ptr(J), _ptr(detJ), _ptr(K), 1, rows, in_vs, gdim, tdim, None,
                             C.byref(ovs)))
        out = np.empty((nc, rows, ovs.value), dtype=np.float32)
        _check(L.bxref32_map(self._h, pull, _ptr(U), _ptr(J), _ptr(detJ), _ptr(K), nc, rows, in_vs, gdim, tdim,
                             _ptr(out), C.byref(ovs)))
        return out

    def push_forward(self, U, J, detJ, K):
        return self._map(0, U, J, detJ, K)

    def pull_back(self, u, J, detJ, K):
        return self._map(1, u, J, detJ, K)

    def T_apply_batch(self, op: str, data: np.ndarray, n: int, cell_info: np.ndarray):
        assert data.dtype == np.float32 and data.flags.c_contiguous and data.ndim == 2
        ci = np.ascontiguousarray(cell_info, dtype=np.uint32)
        _check(lib().bxref32_T_apply(self._h, T_OPS[op], _ptr(data), data.shape[0], data.shape[1], int(n), _ptr(ci)))
        return data
