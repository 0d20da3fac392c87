"""Host-side mirror of the reference's ``basix.finite_element.FiniteElement``
(python/basix/finite_element.py:50-592) over the CUDA C ABI.

Same method names, argument order and error behaviour as the reference for the
hot-path calls (``tabulate``, ``push_forward``, ``pull_back``, ``T_apply``,
``Tt_apply_right``, ``Tt_inv_apply`` ...), plus batched variants the reference
cannot express (one ``cell_info`` per cell).  Arrays may be

* numpy arrays (host memory): the library stages them through the GPU and
  returns numpy arrays -- the drop-in behaviour;
* torch CUDA tensors (device memory): kernels are enqueued on torch's current
  stream and torch tensors are returned without any host round trip.

All arithmetic happens in ``libbasix_b200.so``; nothing here computes.
"""

from __future__ import annotations

import ctypes as C
from enum import IntEnum

import numpy as np

from . import _capi
from ._capi import BX_MEM_DEVICE, BX_MEM_HOST, ElementDesc, check, lib


class CellType(IntEnum):
    """basix::cell::type, cpp/basix/cell.h:20-30."""

    point = 0
    interval = 1
    triangle = 2
    tetrahedron = 3
    quadrilateral = 4
    hexahedron = 5
    prism = 6
    pyramid = 7


class ElementFamily(IntEnum):
    """basix::element::family, cpp/basix/element-families.h:44-61."""

    custom = 0
    P = 1
    RT = 2
    N1E = 3
    BDM = 4
    N2E = 5
    CR = 6
    Regge = 7
    DPC = 8
    bubble = 9
    serendipity = 10
    HHJ = 11
    Hermite = 12
    iso = 13


class LagrangeVariant(IntEnum):
    """basix::element::lagrange_variant, cpp/basix/element-families.h:12-27."""

    unset = 0
    equispaced = 1
    gll_warped = 2
    gll_isaac = 3
    gll_centroid = 4
    chebyshev_warped = 5
    chebyshev_isaac = 6
    chebyshev_centroid = 7
    gl_warped = 8
    gl_isaac = 9
    gl_centroid = 10
    legendre = 11
    bernstein = 12


class DPCVariant(IntEnum):
    """basix::element::dpc_variant, cpp/basix/element-families.h:32-42."""

    unset = 0
    simplex_equispaced = 1
    simplex_gll = 2
    horizontal_equispaced = 3
    horizontal_gll = 4
    diagonal_equispaced = 5
    diagonal_gll = 6
    legendre = 7


class MapType(IntEnum):
    """basix::maps::type, cpp/basix/maps.h:39-47."""

    identity = 0
    L2Piola = 1
    covariantPiola = 2
    contravariantPiola = 3
    doubleCovariantPiola = 4
    doubleContravariantPiola = 5


class PolysetType(IntEnum):
    """basix::polyset::type, cpp/basix/polyset.h:136-140."""

    standard = 0
    macroedge = 1


_TDIM = {0: 0, 1: 1, 2: 2, 3: 3, 4: 2, 5: 3, 6: 3, 7: 3}
_NUM_ENTITIES = {0: (1,), 1: (2, 1), 2: (3, 3, 1), 3: (4, 6, 4, 1), 4: (4, 4, 1), 5: (8, 12, 6, 1),
                 6: (6, 9, 5, 1), 7: (5, 8, 5, 1)}

# cell::topology (cpp/basix/cell.cpp:52-141): vertices of every sub-entity, per dimension
_TOPOLOGY = {
    0: [[[0]]],
    1: [[[0], [1]], [[0, 1]]],
    2: [[[0], [1], [2]], [[1, 2], [0, 2], [0, 1]], [[0, 1, 2]]],
    3: [[[0], [1], [2], [3]], [[2, 3], [1, 3], [1, 2], [0, 3], [0, 2], [0, 1]],
        [[1, 2, 3], [0, 2, 3], [0, 1, 3], [0, 1, 2]], [[0, 1, 2, 3]]],
    4: [[[0], [1], [2], [3]], [[0, 1], [0, 2], [1, 3], [2, 3]], [[0, 1, 2, 3]]],
    5: [[[i] for i in range(8)],
        [[0, 1], [0, 2], [0, 4], [1, 3], [1, 5], [2, 3], [2, 6], [3, 7], [4, 5], [4, 6], [5, 7], [6, 7]],
        [[0, 1, 2, 3], [0, 1, 4, 5], [0, 2, 4, 6], [1, 3, 5, 7], [2, 3, 6, 7], [4, 5, 6, 7]], [list(range(8))]],
    6: [[[i] for i in range(6)], [[0, 1], [0, 2], [0, 3], [1, 2], [1, 4], [2, 5], [3, 4], [3, 5], [4, 5]],
        [[0, 1, 2], [0, 1, 3, 4], [0, 2, 3, 5], [1, 2, 4, 5], [3, 4, 5]], [list(range(6))]],
    7: [[[i] for i in range(5)], [[0, 1], [0, 2], [0, 4], [1, 3], [1, 4], [2, 3], [2, 4], [3, 4]],
        [[0, 1, 2, 3], [0, 1, 4], [0, 2, 4], [1, 3, 4], [2, 3, 4]], [list(range(5))]],
}

T_OPS = {"T_apply": 0, "Tt_apply": 1, "Tt_inv_apply": 2, "Tinv_apply": 3, "Tt_apply_right": 4,
         "Tinv_apply_right": 5, "T_apply_right": 6, "Tt_inv_apply_right": 7}


# ---- array plumbing ------------------------------------------------------------------------------
def _is_torch(a) -> bool:
    return type(a).__module__.startswith("torch")


class _Arg:
    """A C-contiguous array argument: pointer + memory kind (+ keeps the owner alive)."""

    __slots__ = ("owner", "ptr", "mem", "shape")

    def __init__(self, a, dtype, name: str, writable: bool = False):
        if _is_torch(a):
            import torch

            want = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32,
                    np.uint32: torch.uint32}[dtype]
            if not a.is_cuda:
                raise TypeError(f"{name}: torch tensors must live on a CUDA device (use numpy for host data)")
            if want == torch.uint32 and a.dtype == torch.int32:
                a = a.view(torch.uint32)  # same bits
            if a.dtype != want:
                if writable:
                    raise TypeError(f"{name}: expected dtype {want}, got {a.dtype}")
                a = a.to(want)
            if not a.is_contiguous():
                if writable:
                    raise TypeError(f"{name}: in-place argument must be contiguous")
                a = a.contiguous()
            self.owner, self.ptr, self.mem, self.shape = a, a.data_ptr(), BX_MEM_DEVICE, tuple(a.shape)
        else:
            if writable:
                if not isinstance(a, np.ndarray) or a.dtype != dtype or not a.flags.c_contiguous:
                    # the reference binds in-place arguments with .noconvert() (basix_wrappers.h:135,205)
                    raise TypeError(f"{name}: in-place argument must be a C-contiguous numpy array of {dtype.__name__}")
                if not a.flags.writeable:
                    raise TypeError(f"{name}: array is read-only")
            else:
                a = np.ascontiguousarray(a, dtype=dtype)
            self.owner, self.ptr, self.mem, self.shape = a, a.ctypes.data, BX_MEM_HOST, a.shape


def _same_mem(*args: _Arg) -> int:
    mems = {a.mem for a in args}
    if len(mems) != 1:
        raise TypeError("all array arguments must be of one kind: numpy (host) or torch CUDA tensors (device)")
    return mems.pop()


def _stream_and_like(arg: _Arg):
    """(cuda stream handle, allocator for outputs of the same kind)."""
    if arg.mem == BX_MEM_DEVICE:
        import torch

        dev = arg.owner.device
        stream = torch.cuda.current_stream(dev).cuda_stream

        def alloc(shape, dtype):
            td = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32}[dtype]
            return torch.empty(shape, dtype=td, device=dev)

        return stream, alloc
    return None, lambda shape, dtype: np.empty(shape, dtype=dtype)


def _float_dtype(a):
    """np.float32 when ``a`` is a float32 numpy array / torch tensor, else np.float64 (everything else is
    converted to float64, as numpy callers of the reference's float64 element expect)."""
    if _is_torch(a):
        import torch

        return np.float32 if a.dtype == torch.float32 else np.float64
    return np.float32 if getattr(a, "dtype", None) == np.float32 else np.float64


def _ptr_of(a) -> int:
    return a.data_ptr() if _is_torch(a) else a.ctypes.data


# ---- polynomial sets -------------------------------------------------------------------------------
def polyset_dim(celltype: int, ptype: int, degree: int) -> int:
    return lib().bx_polyset_dim(int(celltype), int(ptype), int(degree))


def polyset_nderivs(celltype: int, nderiv: int) -> int:
    return lib().bx_polyset_nderivs(int(celltype), int(nderiv))


def tabulate_polynomial_set(celltype: int, ptype: int, degree: int, nderiv: int, x):
    """``basix._basixcpp.tabulate_polynomial_set_float64`` (basix_wrappers.h:467-475):
    returns P with shape (nderivs, psize, npts)."""
    tdim = _TDIM[int(celltype)]
    xa = _Arg(x, np.float64, "x")
    shape = xa.shape
    if len(shape) == 1:
        shape = (shape[0], 1)
    if len(shape) != 2 or (tdim > 0 and shape[1] != tdim):
        raise RuntimeError(f"Point dim ({shape[1] if len(shape) == 2 else '?'}) does not match cell dim ({tdim}).")
    stream, alloc = _stream_and_like(xa)
    P = alloc((polyset_nderivs(celltype, nderiv), polyset_dim(celltype, ptype, degree), shape[0]), np.float64)
    check(lib().bx_polyset_tabulate_f64(int(celltype), int(ptype), int(degree), int(nderiv), xa.ptr, shape[0],
                                        _ptr_of(P), xa.mem, stream))
    return P


# ---- cell geometry -----------------------------------------------------------------------------------
def geometry(coords, dphi=None, tdim: int | None = None, want=("J", "detJ", "K")):
    """J, detJ and K of every cell (``bx_geometry_f64`` / ``_f32``): what callers compute before ``push_forward`` /
    ``pull_back``; the reference's own analogue is the affine case ``cell::entity_jacobians`` (cell.cpp:662-689).

    ``coords``: (ncells, nnodes, gdim).  ``dphi``: (tdim, npts, nnodes), the derivative slabs ``tabulate(1, X)[1:, :, :, 0]``
    of the scalar coordinate element; ``None`` for affine simplex cells (nnodes = tdim + 1, one point:
    ``J[:, t] = x[1 + t] - x[0]``).  Returns ``(J, detJ, K)`` with shapes (ncells, npts, gdim, tdim), (ncells, npts),
    (ncells, npts, tdim, gdim) -- for ``dphi=None`` without the point axis, ready for ``push_forward`` -- or ``None`` for
    the arrays not named in ``want``.  numpy in -> numpy out, CUDA torch tensors in -> CUDA torch tensors out."""
    dt = _float_dtype(coords)
    ca = _Arg(coords, dt, "coords")
    if len(ca.shape) != 3:
        raise TypeError("coords must have shape (ncells, nnodes, gdim)")
    nc, nv, gdim = ca.shape
    if dphi is None:
        td, npts, dptr = (nv - 1 if tdim is None else int(tdim)), 1, None
        mem = ca.mem
    else:
        da = _Arg(dphi, dt, "dphi")
        if len(da.shape) != 3 or da.shape[2] != nv:
            raise TypeError("dphi must have shape (tdim, npts, nnodes)")
        td, npts, dptr = da.shape[0], da.shape[1], da.ptr
        mem = _same_mem(ca, da)
    stream, alloc = _stream_and_like(ca)
    lead = (nc,) if dphi is None else (nc, npts)
    J = alloc(lead + (gdim, td), dt) if "J" in want else None
    detJ = alloc(lead, dt) if "detJ" in want else None
    K = alloc(lead + (td, gdim), dt) if "K" in want else None
    fn = lib().bx_geometry_f32 if dt is np.float32 else lib().bx_geometry_f64
    check(fn(ca.ptr, dptr, nc, nv, gdim, td, npts, None if J is None else _ptr_of(J), None if detJ is None else _ptr_of(detJ),
             None if K is None else _ptr_of(K), mem, stream))
    return J, detJ, K


# ---- element ----------------------------------------------------------------------------------------
class FiniteElement:
    """B200-backed finite element: the hot-path subset of ``basix.finite_element.FiniteElement``."""

    def __init__(self, *, cell_type, polyset_type, degree, embedded_superdegree, map_type, value_shape,
                 coefficient_matrix, entity_dofs, entity_transformations, dof_ordering=None, family=0,
                 lagrange_variant=0, dpc_variant=0, discontinuous=False, points=None, interpolation_matrix=None):
        self._h = None
        self._cell_type = CellType(int(cell_type))
        self._polyset_type = PolysetType(int(polyset_type))
        self._degree = int(degree)
        self._embedded_superdegree = int(embedded_superdegree)
        self._map_type = MapType(int(map_type))
        self._value_shape = tuple(int(s) for s in value_shape)
        self._family = ElementFamily(int(family))
        self._lagrange_variant = LagrangeVariant(int(lagrange_variant))
        self._dpc_variant = DPCVariant(int(dpc_variant))
        self._discontinuous = bool(discontinuous)
        vs = int(np.prod(self._value_shape)) if self._value_shape else 1
        coeffs = np.ascontiguousarray(coefficient_matrix, dtype=np.float64)
        if coeffs.ndim != 2:
            raise ValueError("coefficient_matrix must be 2-D")
        self._coeffs = coeffs
        tdim = _TDIM[int(cell_type)]
        counts = _NUM_ENTITIES[int(cell_type)]
        if len(entity_dofs) != tdim + 1 or any(len(entity_dofs[d]) != counts[d] for d in range(tdim + 1)):
            raise ValueError("entity_dofs does not match the cell topology")
        self._entity_dofs = [[[int(v) for v in e] for e in row] for row in entity_dofs]
        flat, offs = [], [0]
        for row in self._entity_dofs:
            for e in row:
                flat.extend(e)
                offs.append(len(flat))
        offs_a = np.asarray(offs, dtype=np.int32)
        flat_a = np.asarray(flat if flat else [0], dtype=np.int32)
        self._entity_transformations = {int(k): np.ascontiguousarray(v, dtype=np.float64)
                                        for k, v in entity_transformations.items()}
        self._dof_ordering = (np.asarray(dof_ordering, dtype=np.int32)
                              if dof_ordering is not None and len(dof_ordering) else np.zeros(0, dtype=np.int32))

        d = ElementDesc()
        d.cell_type, d.polyset_type, d.degree = int(cell_type), int(polyset_type), int(degree)
        d.embedded_superdegree, d.map_type = int(embedded_superdegree), int(map_type)
        d.ndofs, d.value_size = coeffs.shape[0], vs
        d.coeffs = coeffs.ctypes.data
        d.dof_ordering = self._dof_ordering.ctypes.data if len(self._dof_ordering) else None
        d.n_dof_ordering = len(self._dof_ordering)
        d.entity_dof_offsets, d.entity_dofs = offs_a.ctypes.data, flat_a.ctypes.data
        keep = []
        for ct, name in ((1, "interval"), (2, "triangle"), (4, "quadrilateral")):
            m = self._entity_transformations.get(ct)
            if m is not None:
                if m.ndim != 3 or m.shape[1] != m.shape[2] or m.shape[0] != (1 if ct == 1 else 2):
                    raise ValueError(f"entity_transformations[{name}] has a bad shape {m.shape}")
                keep.append(m)
                # a zero-sized matrix still needs a non-null pointer so "present, n = 0" is expressible
                setattr(d, f"etrans_{name}", m.ctypes.data if m.size else flat_a.ctypes.data)
                setattr(d, f"etrans_{name}_dim", m.shape[1])
        if coeffs.shape[1] != vs * polyset_dim(cell_type, polyset_type, embedded_superdegree):
            raise ValueError("coefficient_matrix has the wrong number of columns for this polyset")
        # optional: points() and interpolation_matrix() (needed by interpolate_batch only)
        self._points = self._interpolation_matrix = None
        if points is not None and interpolation_matrix is not None:
            self._points = np.ascontiguousarray(points, dtype=np.float64)
            self._interpolation_matrix = np.ascontiguousarray(interpolation_matrix, dtype=np.float64)
            npts = self._points.shape[0]
            if self._points.ndim != 2 or self._points.shape[1] != tdim or \
                    self._interpolation_matrix.shape != (coeffs.shape[0], vs * npts):
                raise ValueError("points must be (npts, tdim) and interpolation_matrix (ndofs, value_size * npts)")
            d.points, d.npts = self._points.ctypes.data, npts
            d.interpolation_matrix = self._interpolation_matrix.ctypes.data
        h = C.c_void_p()
        check(lib().bx_element_create_from_arrays(C.byref(d), C.byref(h)))
        self._h = h
        info = (C.c_int32 * 10)()
        check(lib().bx_element_info(self._h, info))
        self._tdim, self._ndofs, self._vs = info[2], info[3], info[4]
        self._perms, self._identity, self._psize = bool(info[7]), bool(info[8]), info[9]

    @classmethod
    def _from_handle(cls, h, *, family, lagrange_variant, dpc_variant, discontinuous, degree):
        """Wrap an element built inside the library (``bx_create_element``): read the defining arrays back."""
        self = cls.__new__(cls)
        self._h = h
        L = lib()
        info = (C.c_int32 * 10)()
        check(L.bx_element_info(h, info))
        self._cell_type, self._polyset_type = CellType(info[0]), PolysetType(info[1])
        self._tdim, self._ndofs, self._vs = info[2], info[3], info[4]
        self._map_type, self._embedded_superdegree = MapType(info[5]), info[6]
        self._perms, self._identity, self._psize = bool(info[7]), bool(info[8]), info[9]
        self._degree = int(degree)
        self._family, self._lagrange_variant = ElementFamily(int(family)), LagrangeVariant(int(lagrange_variant))
        self._dpc_variant, self._discontinuous = DPCVariant(int(dpc_variant)), bool(discontinuous)
        self._value_shape = () if self._vs == 1 else (self._vs,)
        self._coeffs = np.empty((self._ndofs, self._vs * self._psize), dtype=np.float64)
        check(L.bx_element_coefficient_matrix(h, self._coeffs.ctypes.data))
        self._entity_dofs = []
        for dim, cnt in enumerate(_NUM_ENTITIES[int(self._cell_type)]):
            row = []
            for i in range(cnt):
                n = L.bx_element_entity_dofs(h, dim, i, None)
                buf = np.zeros(max(n, 1), dtype=np.int32)
                L.bx_element_entity_dofs(h, dim, i, buf.ctypes.data)
                row.append([int(v) for v in buf[:n]])
            self._entity_dofs.append(row)
        self._entity_transformations = {}
        for ct in (1, 2, 4):
            shape = (C.c_int32 * 3)()
            if L.bx_element_entity_transformations(h, ct, shape, None):
                m = np.empty(tuple(shape), dtype=np.float64)
                L.bx_element_entity_transformations(h, ct, shape, m.ctypes.data if m.size else None)
                self._entity_transformations[ct] = m
        n = L.bx_element_dof_ordering(h, None)
        self._dof_ordering = np.zeros(n, dtype=np.int32)
        if n:
            L.bx_element_dof_ordering(h, self._dof_ordering.ctypes.data)
        ishape = (C.c_int32 * 4)()
        check(L.bx_element_interpolation_shape(h, ishape))
        self._points = self._interpolation_matrix = None
        if ishape[0] > 0:
            self._points = np.empty((ishape[0], ishape[1]), dtype=np.float64)
            self._interpolation_matrix = np.empty((ishape[2], ishape[3]), dtype=np.float64)
            check(L.bx_element_points(h, self._points.ctypes.data))
            check(L.bx_element_interpolation_matrix(h, self._interpolation_matrix.ctypes.data))
        return self

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            try:
                lib().bx_element_destroy(h)
            except Exception:
                pass

    # ---- serialisation of the defining arrays (fixtures, caches) ---------------------------------
    def to_arrays(self) -> dict:
        out = {
            "meta": np.asarray([int(self._cell_type), int(self._polyset_type), self._degree,
                                self._embedded_superdegree, int(self._map_type), int(self._family),
                                int(self._lagrange_variant), int(self._dpc_variant), int(self._discontinuous)],
                               dtype=np.int64),
            "value_shape": np.asarray(self._value_shape, dtype=np.int64),
            "coefficient_matrix": self._coeffs,
            "dof_ordering": self._dof_ordering,
        }
        flat, offs = [], [0]
        for row in self._entity_dofs:
            for e in row:
                flat.extend(e)
                offs.append(len(flat))
        out["entity_dof_offsets"] = np.asarray(offs, dtype=np.int32)
        out["entity_dofs_flat"] = np.asarray(flat, dtype=np.int32)
        for ct, m in self._entity_transformations.items():
            out[f"etrans_{ct}"] = m
        return out

    @classmethod
    def from_arrays(cls, arrs) -> "FiniteElement":
        meta = [int(v) for v in arrs["meta"]]
        cell = meta[0]
        counts = _NUM_ENTITIES[cell]
        offs, flat = arrs["entity_dof_offsets"], arrs["entity_dofs_flat"]
        edofs, k = [], 0
        for n in counts:
            row = []
            for _ in range(n):
                row.append([int(v) for v in flat[offs[k]:offs[k + 1]]])
                k += 1
            edofs.append(row)
        etrans = {int(key.split("_")[1]): arrs[key] for key in arrs.keys() if key.startswith("etrans_")}
        keys = set(arrs.keys())
        extra = ({"points": arrs["points"], "interpolation_matrix": arrs["interpolation_matrix"]}
                 if {"points", "interpolation_matrix"} <= keys else {})
        return cls(**extra, cell_type=cell, polyset_type=meta[1], degree=meta[2], embedded_superdegree=meta[3],
                   map_type=meta[4], family=meta[5], lagrange_variant=meta[6], dpc_variant=meta[7],
                   discontinuous=bool(meta[8]), value_shape=tuple(int(v) for v in arrs["value_shape"]),
                   coefficient_matrix=arrs["coefficient_matrix"], entity_dofs=edofs,
                   entity_transformations=etrans, dof_ordering=arrs["dof_ordering"])

    # ---- properties (finite_element.py:407-592 of the reference) --------------------------------------
    @property
    def cell_type(self) -> CellType:
        return self._cell_type

    @property
    def polyset_type(self) -> PolysetType:
        return self._polyset_type

    @property
    def degree(self) -> int:
        return self._degree

    @property
    def embedded_superdegree(self) -> int:
        return self._embedded_superdegree

    @property
    def family(self) -> ElementFamily:
        return self._family

    @property
    def lagrange_variant(self) -> LagrangeVariant:
        return self._lagrange_variant

    @property
    def dpc_variant(self) -> DPCVariant:
        return self._dpc_variant

    @property
    def discontinuous(self) -> bool:
        return self._discontinuous

    @property
    def dim(self) -> int:
        return self._ndofs

    @property
    def value_size(self) -> int:
        return self._vs

    @property
    def value_shape(self) -> tuple:
        return self._value_shape

    @property
    def map_type(self) -> MapType:
        return self._map_type

    @property
    def dof_transformations_are_permutations(self) -> bool:
        return self._perms

    @property
    def dof_transformations_are_identity(self) -> bool:
        return self._identity

    @property
    def coefficient_matrix(self) -> np.ndarray:
        return self._coeffs

    @property
    def entity_dofs(self):
        return self._entity_dofs

    @property
    def dof_ordering(self):
        return [int(v) for v in self._dof_ordering]

    # ---- element metadata derived from the defining arrays (set-up side of the reference, host only) -------
    @property
    def dtype(self):
        """``FiniteElement.dtype``: the element object serves float64 and float32 arrays alike; float64 is the type
        of its own tables."""
        return np.float64

    @property
    def num_entity_dofs(self):
        """``FiniteElement.num_entity_dofs`` (finite-element.h:658-663)."""
        return [[len(e) for e in row] for row in self._entity_dofs]

    @property
    def entity_closure_dofs(self):
        """``FiniteElement.entity_closure_dofs`` (finite-element.h:686-690; built at finite-element.cpp:1238-1256): the
        dofs of every sub-entity contained in the entity, sorted."""
        topo = _TOPOLOGY[int(self._cell_type)]
        out = []
        for d, ents in enumerate(topo):
            row = []
            for verts in ents:
                vs = set(verts)
                dofs = [q for dim in range(d + 1) for i, sub in enumerate(topo[dim]) if set(sub) <= vs
                        for q in self._entity_dofs[dim][i]]
                row.append(sorted(dofs))
            out.append(row)
        return out

    @property
    def num_entity_closure_dofs(self):
        """``FiniteElement.num_entity_closure_dofs`` (finite-element.h:673-678)."""
        return [[len(e) for e in row] for row in self.entity_closure_dofs]

    @property
    def interpolation_is_identity(self) -> bool:
        """``FiniteElement.interpolation_is_identity`` (finite-element.cpp:1707-1722); None without interpolation data."""
        M = self._interpolation_matrix
        if M is None:
            return None
        return bool(M.shape[0] == M.shape[1] and np.all(np.abs(M - np.eye(M.shape[0])) <= 100 * np.finfo(np.float64).eps))

    @property
    def interpolation_nderivs(self) -> int:
        """``FiniteElement.interpolation_nderivs``: 0 for every family built here (point evaluations and integral
        moments of values)."""
        return 0

    @property
    def embedded_subdegree(self) -> int:
        """``FiniteElement.embedded_subdegree`` for the natively built families (e-lagrange.cpp:660-664: degree;
        e-raviart-thomas.cpp:160-164, e-nedelec.cpp:342-346: degree - 1)."""
        fam = int(self._family)
        if fam == 1:
            return self._degree
        if fam in (2, 3):
            return self._degree - 1
        raise RuntimeError("embedded_subdegree: only known for the families built natively (P, RT, N1E)")

    @property
    def sobolev_space(self) -> int:
        """``FiniteElement.sobolev_space`` as the reference's enum value (sobolev-spaces.h:12-22): L2 = 0 for
        discontinuous elements, else H1 = 1 (P), HDiv = 10 (RT), HCurl = 11 (N1E)."""
        if self._discontinuous:
            return 0
        fam = int(self._family)
        if fam in (1, 2, 3):
            return {1: 1, 2: 10, 3: 11}[fam]
        raise RuntimeError("sobolev_space: only known for the families built natively (P, RT, N1E)")

    def base_transformations(self) -> np.ndarray:
        """``FiniteElement.base_transformations`` (finite-element.cpp:1902-1969): one (ndofs, ndofs) matrix per edge
        (reversal) and two per face (rotation, reflection), the entity's block placed on the running dof offset."""
        tdim, n = self._tdim, self._ndofs
        mats = []
        start = sum(len(v) for v in self._entity_dofs[0]) if tdim > 0 else 0
        if tdim > 1:
            t = self._entity_transformations[1]
            for e in self._entity_dofs[1]:
                k = len(e)
                m = np.eye(n)
                m[start:start + k, start:start + k] = t[0, :k, :k]
                mats.append(m)
                start += k
        if tdim > 2:
            face_type = {3: lambda f: 2, 5: lambda f: 4, 6: lambda f: 2 if f in (0, 4) else 4, 7: lambda f: 4 if f == 0 else 2}
            for f, e in enumerate(self._entity_dofs[2]):
                k = len(e)
                if k == 0:
                    continue
                t = self._entity_transformations[face_type[int(self._cell_type)](f)]
                for i in range(2):
                    m = np.eye(n)
                    m[start:start + k, start:start + k] = t[i, :k, :k]
                    mats.append(m)
                start += k
        # num_transformations(cell) matrices are returned; faces without dofs leave identities at the END of the
        # array (the reference does not advance its counter for them, finite-element.cpp:1948-1963)
        nt = (len(self._entity_dofs[1]) if tdim > 1 else 0) + (2 * len(self._entity_dofs[2]) if tdim > 2 else 0)
        mats += [np.eye(n)] * (nt - len(mats))
        return np.array(mats).reshape(nt, n, n)

    def __eq__(self, other) -> bool:
        """``FiniteElement.__eq__`` (finite-element.cpp:1727-1764): same defining data."""
        if not isinstance(other, FiniteElement):
            return NotImplemented
        return (self._cell_type == other._cell_type and self._polyset_type == other._polyset_type
                and self._map_type == other._map_type and self._value_shape == other._value_shape
                and self._embedded_superdegree == other._embedded_superdegree
                and self._entity_dofs == other._entity_dofs and self.dof_ordering == other.dof_ordering
                and self._coeffs.shape == other._coeffs.shape and bool(np.allclose(self._coeffs, other._coeffs, rtol=0, atol=1e-12)))

    def __hash__(self) -> int:
        return hash((int(self._cell_type), int(self._map_type), self._value_shape, self._embedded_superdegree, self._ndofs,
                     tuple(self.dof_ordering)))

    def entity_transformations(self) -> dict:
        names = {1: "interval", 2: "triangle", 4: "quadrilateral"}
        return {names[k]: v for k, v in self._entity_transformations.items()}

    def net_permutation(self, cell_info: int, op: str = "T_apply") -> np.ndarray:
        """Host tables: src with out[i] = in[src[i]] for this cell_info (permutation elements)."""
        src = np.empty(self._ndofs, dtype=np.int32)
        check(lib().bx_element_net_operator(self._h, T_OPS[op], int(cell_info) & 0xFFFFFFFF, src.ctypes.data, None))
        return src

    def net_operator(self, cell_info: int, op: str = "T_apply") -> np.ndarray:
        """Host tables: dense (ndofs, ndofs) matrix A with out = A @ in for this cell_info."""
        A = np.empty((self._ndofs, self._ndofs), dtype=np.float64)
        check(lib().bx_element_net_operator(self._h, T_OPS[op], int(cell_info) & 0xFFFFFFFF, None, A.ctypes.data))
        return A

    def tabulate_shape(self, nd: int, npts: int):
        shape = (C.c_size_t * 4)()
        check(lib().bx_element_tabulate_shape(self._h, int(nd), int(npts), shape))
        return tuple(int(s) for s in shape)

    def prepare(self, nd: int) -> None:
        """Build and upload the tabulate operands for derivative orders 0..nd now (``bx_element_prepare``): afterwards
        device-pointer ``tabulate`` calls only launch kernels (safe to capture into a CUDA graph).  Done at creation for
        nd <= 2."""
        check(lib().bx_element_prepare(self._h, int(nd)))

    def tabulate_path(self, nd: int) -> str:
        return {0: "generic", 1: "fused", 2: "tensor-product", 3: "register", 4: "derivative-operator"}.get(lib().bx_element_tabulate_path(self._h, int(nd)), "?")

    # ---- hot path ---------------------------------------------------------------------------------------
    def tabulate(self, n: int, x, out=None):
        """``FiniteElement.tabulate`` (finite_element.py:63-95): shape (nderivs, npts, ndofs, value_size).

        float64 or float32 points (the dtype of ``x`` selects ``bx_tabulate_f64`` / ``_f32``, as the reference's
        ``FiniteElement_float64`` / ``_float32`` classes do).  ``out`` (optional, same kind and dtype as ``x``)
        mirrors the reference's C++ overload with the basis data as an out argument (finite-element.h:454-455),
        avoiding the allocation on repeated calls."""
        dt = _float_dtype(x)
        xa = _Arg(x, dt, "x")
        if len(xa.shape) != 2:
            raise TypeError("x must be a 2-D array of shape (npts, tdim)")
        npts, xdim = xa.shape
        stream, alloc = _stream_and_like(xa)
        if xdim != self._tdim:
            raise RuntimeError(f"Point dim ({xdim}) does not match element dim ({self._tdim}).")
        shape = self.tabulate_shape(n, npts)
        if out is None:
            out = alloc(shape, dt)
        else:
            oa = _Arg(out, dt, "out", writable=True)
            if oa.mem != xa.mem or tuple(oa.shape) != shape:
                raise RuntimeError(f"out must have shape {shape} and live where x lives")
        fn = lib().bx_tabulate_f32 if dt is np.float32 else lib().bx_tabulate_f64
        check(fn(self._h, int(n), xa.ptr, npts, xdim, _ptr_of(out), xa.mem, stream))
        return out

    def push_forward_broadcast(self, U, J, detJ, K):
        """``push_forward`` of one slab of reference values ``U`` (rows, value_size) onto every cell: same result as
        ``push_forward(np.broadcast_to(U, (ncells, rows, value_size)), J, detJ, K)`` without replicating ``U`` (the
        tabulated reference basis is the same on all cells).  Returns (ncells, rows, physical value size)."""
        dt = _float_dtype(J)
        Ua, Ja = _Arg(U, dt, "U"), _Arg(J, dt, "J")
        da, Ka = _Arg(detJ, dt, "detJ"), _Arg(K, dt, "K")
        mem = _same_mem(Ua, Ja, da, Ka)
        if len(Ua.shape) != 2 or len(Ja.shape) != 3 or len(Ka.shape) != 3 or len(da.shape) != 1:
            raise TypeError("expected U (rows, vs), J (nc, gdim, tdim), detJ (nc,), K (nc, tdim, gdim)")
        rows, in_vs = Ua.shape
        nc, gdim, tdim = Ja.shape
        if da.shape[0] != nc or Ka.shape != (nc, tdim, gdim):
            raise RuntimeError("push_forward: inconsistent array shapes")
        out_vs = lib().bx_map_value_size(int(self._map_type), 0, in_vs, gdim, tdim)
        if out_vs < 0:
            check(1)
        stream, alloc = _stream_and_like(Ja)
        out = alloc((nc, rows, out_vs), dt)
        fn = lib().bx_push_forward_broadcast_f32 if dt is np.float32 else lib().bx_push_forward_broadcast_f64
        check(fn(self._h, Ua.ptr, Ja.ptr, da.ptr, Ka.ptr, nc, rows, in_vs, gdim, tdim, _ptr_of(out), mem, stream))
        return out

    def physical_basis(self, X, coords=None, J=None, detJ=None, K=None, cell_info=None, op: str = "T_apply", U=None):
        """Basis functions of every physical cell in one fused pass (``bx_physical_basis_f64`` /
        ``bx_push_forward_fused_*``): ``out[c, p] = push_forward(T_op(cell_info[c]) tabulate(0, X)[0][p], J_c, detJ_c, K_c)``
        -- the reference's three calls (``tabulate`` once; per cell ``T_apply`` on the reference values with block size =
        value size, then ``push_forward``).  Geometry: ``coords`` (ncells, tdim + 1, gdim) of affine triangle / tetrahedron
        cells (J, detJ, K are computed on chip), or ``J`` (ncells, gdim, tdim), ``detJ`` (ncells,), ``K`` (ncells, tdim,
        gdim) per cell -- or, for non-affine cells, per (cell, point): ``J`` (ncells, npts, gdim, tdim), ``detJ`` (ncells,
        npts), ``K`` (ncells, npts, tdim, gdim), what ``geometry(coords, dphi)`` returns.  ``cell_info``: (ncells,) uint32 or
        None.  Pass ``U`` = ``tabulate(0, X)[0]`` (npts, ndofs, value_size) instead of ``X`` to reuse a table.  Returns
        (ncells, npts, ndofs, physical value size)."""
        given = U if X is None else X
        dt = _float_dtype(given)
        pointwise = False
        geo = [a for a in (coords, J, detJ, K) if a is not None]
        if not geo and int(self._map_type) != 0:
            raise TypeError("physical_basis: pass coords, or J / detJ / K")
        args = {}
        for name, a in (("coords", coords), ("J", J), ("detJ", detJ), ("K", K)):
            args[name] = None if a is None else _Arg(a, dt, name)
        ci = None if cell_info is None else _Arg(cell_info, np.uint32, "cell_info")
        live = [a for a in args.values() if a is not None] + ([ci] if ci is not None else [])
        ga = _Arg(given, dt, "X" if U is None else "U")
        mem = _same_mem(ga, *live) if live else ga.mem
        if args["coords"] is not None:
            if len(args["coords"].shape) != 3 or args["coords"].shape[1] != self._tdim + 1:
                raise TypeError("coords must have shape (ncells, tdim + 1, gdim)")
            nc, gdim = args["coords"].shape[0], args["coords"].shape[2]
        elif args["K"] is not None or args["J"] is not None:
            m = args["K"] if args["K"] is not None else args["J"]
            if len(m.shape) == 4:
                pointwise = True  # non-affine cells: one Jacobian per (cell, point), as bx.geometry(coords, dphi) returns
            elif len(m.shape) != 3:
                raise TypeError("J must have shape (ncells, [npts,] gdim, tdim) and K (ncells, [npts,] tdim, gdim)")
            nc = m.shape[0]
            gdim = m.shape[-1] if args["K"] is not None else m.shape[-2]
        else:
            if ci is None:
                raise TypeError("physical_basis: the number of cells is unknown (no geometry, no cell_info)")
            nc, gdim = ci.shape[0], self._tdim
        if ci is not None and ci.shape != (nc,):
            raise RuntimeError("cell_info must have one entry per cell")
        out_vs = self._vs if int(self._map_type) == 0 else gdim
        stream, alloc = _stream_and_like(ga)
        ptr = {k: (None if a is None else a.ptr) for k, a in args.items()}
        if pointwise:
            npts = ga.shape[0]
            for name in ("J", "K"):
                if args[name] is not None and (len(args[name].shape) != 4 or args[name].shape[:2] != (nc, npts)):
                    raise RuntimeError("per-point geometry: J (ncells, npts, gdim, tdim), detJ (ncells, npts), K (ncells, npts, "
                                       "tdim, gdim) with npts = the number of points")
            if args["detJ"] is not None and args["detJ"].shape != (nc, npts):
                raise RuntimeError("per-point geometry: detJ must have shape (ncells, npts)")
            out = alloc((nc, npts, self._ndofs, out_vs), dt)
            cip = None if ci is None else ci.ptr
            if dt is np.float32:
                Ut = self.tabulate(0, X)[0] if U is None else U
                ta = _Arg(Ut, dt, "U")
                check(lib().bx_push_forward_fused_pointwise_f32(self._h, T_OPS[op], ta.ptr, npts, ptr["J"], ptr["detJ"], ptr["K"],
                                                                cip, nc, gdim, _ptr_of(out), mem, stream))
                return out
            if U is None:
                if len(ga.shape) != 2:
                    raise TypeError("X must have shape (npts, tdim)")
                check(lib().bx_physical_basis_pointwise_f64(self._h, T_OPS[op], ga.ptr, npts, ga.shape[1], ptr["J"], ptr["detJ"],
                                                            ptr["K"], cip, nc, gdim, _ptr_of(out), mem, stream))
            else:
                if len(ga.shape) != 3 or ga.shape[1:] != (self._ndofs, self._vs):
                    raise TypeError("U must have shape (npts, ndofs, value_size)")
                check(lib().bx_push_forward_fused_pointwise_f64(self._h, T_OPS[op], ga.ptr, npts, ptr["J"], ptr["detJ"], ptr["K"],
                                                                cip, nc, gdim, _ptr_of(out), mem, stream))
            return out
        if U is None:
            if len(ga.shape) != 2:
                raise TypeError("X must have shape (npts, tdim)")
            npts = ga.shape[0]
            out = alloc((nc, npts, self._ndofs, out_vs), dt)
            if dt is np.float32:
                Ut = self.tabulate(0, X)[0]
                check(lib().bx_push_forward_fused_f32(self._h, T_OPS[op], _ptr_of(Ut), npts, ptr["coords"], ptr["J"],
                                                      ptr["detJ"], ptr["K"], None if ci is None else ci.ptr, nc, gdim,
                                                      _ptr_of(out), mem, stream))
            else:
                check(lib().bx_physical_basis_f64(self._h, T_OPS[op], ga.ptr, npts, ga.shape[1], ptr["coords"], ptr["J"],
                                                  ptr["detJ"], ptr["K"], None if ci is None else ci.ptr, nc, gdim,
                                                  _ptr_of(out), mem, stream))
            return out
        if len(ga.shape) != 3 or ga.shape[1:] != (self._ndofs, self._vs):
            raise TypeError("U must have shape (npts, ndofs, value_size)")
        npts = ga.shape[0]
        out = alloc((nc, npts, self._ndofs, out_vs), dt)
        fn = lib().bx_push_forward_fused_f32 if dt is np.float32 else lib().bx_push_forward_fused_f64
        check(fn(self._h, T_OPS[op], ga.ptr, npts, ptr["coords"], ptr["J"], ptr["detJ"], ptr["K"],
                 None if ci is None else ci.ptr, nc, gdim, _ptr_of(out), mem, stream))
        return out

    def _map(self, pull: int, U, J, detJ, K):
        dt = _float_dtype(U)
        Ua, Ja = _Arg(U, dt, "U"), _Arg(J, dt, "J")
        da, Ka = _Arg(detJ, dt, "detJ"), _Arg(K, dt, "K")
        mem = _same_mem(Ua, Ja, da, Ka)
        if len(Ua.shape) != 3 or len(Ja.shape) != 3 or len(Ka.shape) != 3 or len(da.shape) != 1:
            raise TypeError("expected U (nc, rows, vs), J (nc, gdim, tdim), detJ (nc,), K (nc, tdim, gdim)")
        nc, rows, in_vs = Ua.shape
        gdim, tdim = Ja.shape[1], Ja.shape[2]
        if Ja.shape[0] != nc or da.shape[0] != nc or Ka.shape != (nc, tdim, gdim):
            raise RuntimeError("push_forward/pull_back: inconsistent array shapes")
        out_vs = lib().bx_map_value_size(int(self._map_type), pull, in_vs, gdim, tdim)
        if out_vs < 0:
            check(1)
        stream, alloc = _stream_and_like(Ua)
        out = alloc((nc, rows, out_vs), dt)
        L = lib()
        if dt is np.float32:
            fn = L.bx_pull_back_f32 if pull else L.bx_push_forward_f32
        else:
            fn = L.bx_pull_back_f64 if pull else L.bx_push_forward_f64
        check(fn(self._h, Ua.ptr, Ja.ptr, da.ptr, Ka.ptr, nc, rows, in_vs, gdim, tdim, _ptr_of(out), mem, stream))
        return out

    def push_forward(self, U, J, detJ, K):
        """``FiniteElement.push_forward`` (finite_element.py:112-134)."""
        return self._map(0, U, J, detJ, K)

    def pull_back(self, u, J, detJ, K):
        """``FiniteElement.pull_back`` (finite_element.py:136-151)."""
        return self._map(1, u, J, detJ, K)

    def _transform(self, op: str, data, block_size: int, cell_info, batched: bool):
        if _is_torch(data):
            dt = {"torch.float64": np.float64, "torch.float32": np.float32, "torch.int32": np.int32}.get(str(data.dtype))
        else:
            dt = {np.dtype(np.float64): np.float64, np.dtype(np.float32): np.float32,
                  np.dtype(np.int32): np.int32}.get(getattr(data, "dtype", None))
        if dt is None:
            raise TypeError("data must be float64, float32 or int32")
        da = _Arg(data, dt, "data", writable=True)
        if batched:
            if len(da.shape) != 2:
                raise TypeError("batched data must have shape (ncells, ndofs * block_size)")
            nc, cs = da.shape
            ca = _Arg(cell_info, np.uint32, "cell_info")
            if ca.shape != (nc,):
                raise RuntimeError("cell_info must have one entry per cell")
            mem = _same_mem(da, ca)
        else:
            nc, cs = 1, int(np.prod(da.shape))
            if da.mem == BX_MEM_DEVICE:
                import torch

                ca = _Arg(torch.tensor([int(cell_info) & 0xFFFFFFFF], dtype=torch.int64, device=data.device)
                          .to(torch.uint32), np.uint32, "cell_info")
            else:
                ca = _Arg(np.asarray([int(cell_info) & 0xFFFFFFFF], dtype=np.uint32), np.uint32, "cell_info")
            mem = da.mem
        stream, _ = _stream_and_like(da)
        fn = {np.float64: lib().bx_T_apply_f64, np.float32: lib().bx_T_apply_f32, np.int32: lib().bx_T_apply_i32}[dt]
        check(fn(self._h, T_OPS[op], da.ptr, nc, cs, int(block_size), ca.ptr, mem, stream))

    # single-cell calls, reference signatures (finite_element.py:153-194; finite-element.h:1067-1159)
    def T_apply(self, data, block_size: int, cell_info: int) -> None:
        self._transform("T_apply", data, block_size, cell_info, False)

    def Tt_apply(self, data, block_size: int, cell_info: int) -> None:
        self._transform("Tt_apply", data, block_size, cell_info, False)

    def Tt_inv_apply(self, data, block_size: int, cell_info: int) -> None:
        self._transform("Tt_inv_apply", data, block_size, cell_info, False)

    def Tinv_apply(self, data, block_size: int, cell_info: int) -> None:
        self._transform("Tinv_apply", data, block_size, cell_info, False)

    def Tt_apply_right(self, data, block_size: int, cell_info: int) -> None:
        self._transform("Tt_apply_right", data, block_size, cell_info, False)

    def Tinv_apply_right(self, data, block_size: int, cell_info: int) -> None:
        self._transform("Tinv_apply_right", data, block_size, cell_info, False)

    def T_apply_right(self, data, block_size: int, cell_info: int) -> None:
        self._transform("T_apply_right", data, block_size, cell_info, False)

    def Tt_inv_apply_right(self, data, block_size: int, cell_info: int) -> None:
        self._transform("Tt_inv_apply_right", data, block_size, cell_info, False)

    # batched additions: data[ncells][ndofs * block_size], cell_info[ncells]
    def T_apply_batch(self, data, block_size: int, cell_info, op: str = "T_apply") -> None:
        self._transform(op, data, block_size, cell_info, True)

    def _permute(self, inverse: int, d, cell_info, batched: bool):
        da = _Arg(d, np.int32, "d", writable=True)
        if batched:
            if len(da.shape) != 2 or da.shape[1] != self._ndofs:
                raise TypeError("batched d must have shape (ncells, ndofs)")
            nc = da.shape[0]
            ca = _Arg(cell_info, np.uint32, "cell_info")
            if ca.shape != (nc,):
                raise RuntimeError("cell_info must have one entry per cell")
            mem = _same_mem(da, ca)
        else:
            if int(np.prod(da.shape)) != self._ndofs:
                raise RuntimeError("d must have one entry per dof")
            nc = 1
            ca = _Arg(np.asarray([int(cell_info) & 0xFFFFFFFF], dtype=np.uint32), np.uint32, "cell_info")
            if da.mem == BX_MEM_DEVICE:
                import torch

                ca = _Arg(torch.tensor([int(cell_info) & 0xFFFFFFFF], dtype=torch.int64, device=d.device)
                          .to(torch.uint32), np.uint32, "cell_info")
            mem = da.mem
        stream, _ = _stream_and_like(da)
        check(lib().bx_permute_i32(self._h, inverse, da.ptr, nc, ca.ptr, mem, stream))

    def permute(self, d, cell_info: int) -> None:
        """``FiniteElement::permute`` (finite-element.h:800-812), in place."""
        self._permute(0, d, cell_info, False)

    def permute_inv(self, d, cell_info: int) -> None:
        """``FiniteElement::permute_inv`` (finite-element.h:831-843), in place."""
        self._permute(1, d, cell_info, False)

    def permute_batch(self, d, cell_info, inverse: bool = False) -> None:
        self._permute(int(bool(inverse)), d, cell_info, True)

    # ---- interpolation ---------------------------------------------------------------------------------
    @property
    def points(self):
        """``FiniteElement.points`` (finite-element.h:1167-1170): shape (npts, tdim); None when the element was
        described without interpolation data."""
        return self._points

    @property
    def interpolation_matrix(self):
        """``FiniteElement.interpolation_matrix`` (finite-element.h:1222-1226): shape (ndofs, value_size * npts)."""
        return self._interpolation_matrix

    def interpolate_batch(self, values, layout: str = "component-major"):
        """coefficients[c] = interpolation_matrix @ values[c] for every cell (the product the reference leaves to
        the caller, finite-element.h:1183-1216).  ``values``: (ncells, value_size * npts); layout
        "component-major" = the reference's column order ``values[c][j * npts + p]``, "point-major" =
        ``values[c][p * value_size + j]`` (what ``pull_back`` writes with rows = points, reshaped).  numpy in ->
        numpy out, CUDA torch tensor in -> CUDA torch tensor out."""
        if self._points is None:
            raise RuntimeError("interpolate: the element was built without points() / interpolation_matrix()")
        dt = _float_dtype(values)
        va = _Arg(values, dt, "values")
        k = self._vs * self._points.shape[0]
        if len(va.shape) != 2 or va.shape[1] != k:
            raise RuntimeError(f"values must have shape (ncells, {k})")
        stream, alloc = _stream_and_like(va)
        out = alloc((va.shape[0], self._ndofs), dt)
        lay = {"component-major": 0, "point-major": 1}[layout]
        fn = lib().bx_interpolate_f32 if dt is np.float32 else lib().bx_interpolate_f64
        check(fn(self._h, lay, va.ptr, va.shape[0], _ptr_of(out), va.mem, stream))
        return out

    def interpolate_pulled_batch(self, u, J, detJ, K):
        """coefficients[c] = interpolation_matrix @ pull_back(u[c], J[c], detJ[c], K[c]) in one fused pass
        (``bx_interpolate_pulled_f64``): the two steps of interpolating a physical function into a Piola-mapped element,
        without the pulled-back values ever being written.  ``u``: (ncells, npts, gdim) values at the mapped
        interpolation points (the layout ``pull_back`` takes); ``J`` (ncells, gdim, tdim), ``detJ`` (ncells,), ``K``
        (ncells, tdim, gdim).  Covariant / contravariant Piola elements with gdim == tdim; equal to
        ``interpolate_batch(pull_back(u, J, detJ, K).reshape(ncells, -1), "point-major")`` to rounding (1e-13).
        Returns (ncells, ndofs)."""
        if self._points is None:
            raise RuntimeError("interpolate: the element was built without points() / interpolation_matrix()")
        dt = _float_dtype(u)
        ua, Ja = _Arg(u, dt, "u"), _Arg(J, dt, "J")
        da, Ka = _Arg(detJ, dt, "detJ"), _Arg(K, dt, "K")
        mem = _same_mem(ua, Ja, da, Ka)
        npts = self._points.shape[0]
        if len(ua.shape) != 3 or ua.shape[1] != npts or len(Ja.shape) != 3:
            raise RuntimeError(f"u must have shape (ncells, {npts}, gdim) and J (ncells, gdim, tdim)")
        nc, gdim, tdim = Ja.shape
        if ua.shape[0] != nc or ua.shape[2] != gdim or da.shape != (nc,) or Ka.shape != (nc, tdim, gdim):
            raise RuntimeError("interpolate_pulled: inconsistent array shapes")
        stream, alloc = _stream_and_like(ua)
        out = alloc((nc, self._ndofs), dt)
        fn = lib().bx_interpolate_pulled_f32 if dt is np.float32 else lib().bx_interpolate_pulled_f64
        check(fn(self._h, ua.ptr, Ja.ptr, da.ptr, Ka.ptr, nc, gdim, _ptr_of(out), mem, stream))
        return out

    # ---- sub-entity closure permutations ---------------------------------------------------------------
    def subentity_closure_size(self, entity_type) -> int:
        """Number of dofs on the closure of a sub-entity of this type (0 when the cell has none)."""
        return int(lib().bx_element_subentity_closure_size(self._h, int(entity_type)))

    def subentity_closure_map(self, entity_type, entity_info: int, inverse: bool = False):
        """Host table of one state: ``src`` with ``out[k] = in[src[k]]`` (inspection / tests; no device needed)."""
        m = self.subentity_closure_size(entity_type)
        src = np.zeros(max(m, 1), dtype=np.int32)
        check(lib().bx_element_subentity_closure_map(self._h, int(entity_type), int(bool(inverse)),
                                                     int(entity_info) & 0xFFFFFFFF, src.ctypes.data))
        return src[:m]

    def _closure(self, inverse: int, d, info, entity_type, entity_index):
        da = _Arg(d, np.int32, "indices", writable=True)
        if len(da.shape) != 2:
            raise TypeError("batched indices must have shape (ncells, closure size)")
        nc, m = da.shape
        ia = _Arg(info, np.uint32, "info")
        if ia.shape != (nc,):
            raise RuntimeError("info must have one entry per cell")
        mem = _same_mem(da, ia)
        stream, _ = _stream_and_like(da)
        check(lib().bx_permute_subentity_closure_i32(self._h, inverse, da.ptr, nc, m, ia.ptr, int(entity_type),
                                                     -1 if entity_index is None else int(entity_index), mem, stream))

    def permute_subentity_closure(self, indices, cell_or_entity_info: int, entity_type, entity_index=None):
        """``FiniteElement.permute_subentity_closure`` (python/basix/finite_element.py:322-347;
        finite-element.h:860-985): returns the permuted copy of ``indices``, as the reference wrapper does."""
        out = np.ascontiguousarray(indices, dtype=np.int32).copy().reshape(1, -1)
        self._closure(0, out, np.asarray([int(cell_or_entity_info) & 0xFFFFFFFF], dtype=np.uint32), entity_type,
                      entity_index)
        return out[0]

    def permute_subentity_closure_inv(self, indices, cell_or_entity_info: int, entity_type, entity_index=None):
        """``FiniteElement.permute_subentity_closure_inv`` (python/basix/finite_element.py:349-374)."""
        out = np.ascontiguousarray(indices, dtype=np.int32).copy().reshape(1, -1)
        self._closure(1, out, np.asarray([int(cell_or_entity_info) & 0xFFFFFFFF], dtype=np.uint32), entity_type,
                      entity_index)
        return out[0]

    def permute_subentity_closure_batch(self, d, info, entity_type, entity_index=None, inverse: bool = False) -> None:
        """In place on d[ncells][closure size] (numpy int32 or CUDA torch int32), one ``info`` word per cell."""
        self._closure(int(bool(inverse)), d, info, entity_type, entity_index)


def create_element(family, celltype, degree: int, lagrange_variant=LagrangeVariant.unset,
                   dpc_variant=DPCVariant.unset, discontinuous: bool = False, dof_ordering=None,
                   dtype=np.float64) -> FiniteElement:
    """``basix.create_element`` (python/basix/finite_element.py:595-634), same argument order.

    Built natively by the library (host set-up code in csrc/create.cu):

    * P (Lagrange), variants ``equispaced`` / ``gll_warped`` (``unset`` below degree 3), on interval, triangle,
      tetrahedron, quadrilateral, hexahedron; continuous or discontinuous; optional ``dof_ordering``;
    * RT (Raviart-Thomas) and N1E (Nedelec first kind) on triangle and tetrahedron, moment spaces of variant
      ``legendre`` / ``equispaced`` / ``gll_warped`` (``unset`` resolves as in the reference).

    Other families raise ``RuntimeError``; build those with ``FiniteElement(...)`` / ``FiniteElement.from_arrays``
    from the arrays of an element constructed elsewhere."""
    if np.dtype(dtype) != np.float64:
        raise RuntimeError("create_element: elements are described in float64 (float32 data is accepted by the "
                           "T_apply family and tabulate_polynomial_set)")
    order = np.ascontiguousarray(dof_ordering if dof_ordering is not None else [], dtype=np.int32)
    h = C.c_void_p()
    check(lib().bx_create_element(int(family), int(celltype), int(degree), int(lagrange_variant), int(dpc_variant),
                                  int(bool(discontinuous)), order.ctypes.data if len(order) else None, len(order),
                                  C.byref(h)))
    lv = int(lagrange_variant)
    if lv == 0 and int(family) == 1 and degree < 3:
        lv = int(LagrangeVariant.gll_warped)  # the reference resolves "unset" the same way (e-lagrange.cpp:491-500)
    return FiniteElement._from_handle(h, family=family, lagrange_variant=lv, dpc_variant=dpc_variant,
                                      discontinuous=discontinuous, degree=degree)


def compute_interpolation_operator(e0: FiniteElement, e1: FiniteElement, device=None):
    """``basix.compute_interpolation_operator`` (python/basix/interpolation.py; interpolation.cpp:16-116): the matrix
    mapping coefficients in ``e0`` to the coefficients of the interpolant in ``e1``, composed on the device from
    ``tabulate`` (e0 at e1's points) and the batched interpolation (e1's matrix).  Returns a numpy array, or a CUDA
    torch tensor when ``device`` (a torch device) is given."""
    shape = (C.c_int32 * 2)()
    check(lib().bx_compute_interpolation_operator_f64(e0._h, e1._h, shape, None, BX_MEM_HOST, None))
    if device is None:
        out = np.empty((shape[0], shape[1]), dtype=np.float64)
        check(lib().bx_compute_interpolation_operator_f64(e0._h, e1._h, shape, out.ctypes.data, BX_MEM_HOST, None))
        return out
    import torch

    out = torch.empty((shape[0], shape[1]), dtype=torch.float64, device=device)
    check(lib().bx_compute_interpolation_operator_f64(e0._h, e1._h, shape, out.data_ptr(), BX_MEM_DEVICE,
                                                      torch.cuda.current_stream(out.device).cuda_stream))
    return out
