"""``basix.finite_element`` (python/basix/finite_element.py): ``FiniteElement`` wrapper and ``create_element``.

The wrapper keeps the reference's method and property names and forwards to the ``basix_b200`` element, which runs every
compute call on the GPU through the C ABI.  Set-up data of the reference that the hot path never reads (``x``, ``M``,
``wcoeffs``, the tensor-product factorisation) raises :class:`basix.NotBuiltNatively`."""

import numpy as np

from basix._basixcpp import (CellType, DPCVariant, ElementFamily, LagrangeVariant, MapType, NotBuiltNatively,  # noqa: F401
                             PolysetType)
from basix._basixcpp import create_element as _create_element
from basix.sobolev_spaces import SobolevSpace


class FiniteElement:
    """``basix.finite_element.FiniteElement`` (python/basix/finite_element.py:50-592)."""

    def __init__(self, e, dtype=np.float64, tp_factors=None):
        self._e = e
        self._dtype = np.dtype(dtype)
        self._tp_factors = tp_factors

    def __eq__(self, other):
        return isinstance(other, FiniteElement) and self._e == other._e and self._dtype == other._dtype

    def __hash__(self):
        return hash((self._e, self._dtype.char))

    def _cast(self, a):
        return np.ascontiguousarray(a, dtype=self._dtype)

    # ---- hot path ----
    def tabulate(self, n, x):
        return self._e.tabulate(n, self._cast(x))

    def push_forward(self, U, J, detJ, K):
        return self._e.push_forward(*(self._cast(a) for a in (U, J, detJ, K)))

    def pull_back(self, u, J, detJ, K):
        return self._e.pull_back(*(self._cast(a) for a in (u, J, detJ, K)))

    def T_apply(self, u, n, cell_info):
        self._e.T_apply(u, n, cell_info)

    def Tt_apply(self, u, n, cell_info):
        self._e.Tt_apply(u, n, cell_info)

    def Tt_inv_apply(self, u, n, cell_info):
        self._e.Tt_inv_apply(u, n, cell_info)

    def Tinv_apply(self, u, n, cell_info):
        self._e.Tinv_apply(u, n, cell_info)

    def Tt_apply_right(self, u, n, cell_info):
        self._e.Tt_apply_right(u, n, cell_info)

    def Tinv_apply_right(self, u, n, cell_info):
        self._e.Tinv_apply_right(u, n, cell_info)

    def T_apply_right(self, u, n, cell_info):
        self._e.T_apply_right(u, n, cell_info)

    def Tt_inv_apply_right(self, u, n, cell_info):
        self._e.Tt_inv_apply_right(u, n, cell_info)

    def permute_subentity_closure(self, indices, cell_or_entity_info, entity_type, entity_index=None):
        """python/basix/finite_element.py:322-347: returns the permuted copy."""
        out = np.ascontiguousarray(indices, dtype=np.int32).copy()
        self._e.permute_subentity_closure(out, cell_or_entity_info, entity_type, entity_index)
        return out

    def permute_subentity_closure_inv(self, indices, cell_or_entity_info, entity_type, entity_index=None):
        """python/basix/finite_element.py:349-374."""
        out = np.ascontiguousarray(indices, dtype=np.int32).copy()
        self._e.permute_subentity_closure_inv(out, cell_or_entity_info, entity_type, entity_index)
        return out

    # ---- data ----
    def base_transformations(self):
        return self._e.base_transformations().astype(self._dtype, copy=False)

    def entity_transformations(self):
        return {k: v.astype(self._dtype, copy=False) for k, v in self._e.entity_transformations().items()}

    def get_tensor_product_representation(self):
        """``FiniteElement.get_tensor_product_representation`` (finite-element.h:1271-1282)."""
        if self._tp_factors is None:
            raise RuntimeError("Element has no tensor product representation.")
        return self._tp_factors

    @property
    def dtype(self):
        return self._dtype

    @property
    def degree(self):
        return self._e.degree

    @property
    def embedded_superdegree(self):
        return self._e.embedded_superdegree

    @property
    def embedded_subdegree(self):
        return self._e.embedded_subdegree

    @property
    def cell_type(self):
        return self._e.cell_type

    @property
    def polyset_type(self):
        return self._e.polyset_type

    @property
    def dim(self):
        return self._e.dim

    @property
    def num_entity_dofs(self):
        return self._e.num_entity_dofs

    @property
    def entity_dofs(self):
        return self._e.entity_dofs

    @property
    def num_entity_closure_dofs(self):
        return self._e.num_entity_closure_dofs

    @property
    def entity_closure_dofs(self):
        return self._e.entity_closure_dofs

    @property
    def value_size(self):
        return self._e.value_size

    @property
    def value_shape(self):
        return self._e.value_shape

    @property
    def reference_value_shape(self):
        return self._e.value_shape

    @property
    def discontinuous(self):
        return self._e.discontinuous

    @property
    def family(self):
        return self._e.family

    @property
    def lagrange_variant(self):
        return self._e.lagrange_variant

    @property
    def dpc_variant(self):
        return self._e.dpc_variant

    @property
    def dof_transformations_are_permutations(self):
        return self._e.dof_transformations_are_permutations

    @property
    def dof_transformations_are_identity(self):
        return self._e.dof_transformations_are_identity

    @property
    def interpolation_is_identity(self):
        return self._e.interpolation_is_identity

    @property
    def map_type(self):
        return self._e.map_type

    @property
    def sobolev_space(self):
        return SobolevSpace(self._e.sobolev_space)

    @property
    def points(self):
        return self._e.points.astype(self._dtype, copy=False)

    @property
    def interpolation_matrix(self):
        return self._e.interpolation_matrix.astype(self._dtype, copy=False)

    @property
    def interpolation_nderivs(self):
        return self._e.interpolation_nderivs

    @property
    def dof_ordering(self):
        return self._e.dof_ordering

    @property
    def coefficient_matrix(self):
        return self._e.coefficient_matrix.astype(self._dtype, copy=False)

    @property
    def has_tensor_product_factorisation(self):
        return self._tp_factors is not None

    @property
    def polyset_type_name(self):
        return self._e.polyset_type.name

    def _unbuilt(self, name):
        raise NotBuiltNatively(f"FiniteElement.{name}: set-up data of the reference's factory that the hot path never reads")

    @property
    def x(self):
        self._unbuilt("x")

    @property
    def M(self):
        self._unbuilt("M")

    @property
    def wcoeffs(self):
        self._unbuilt("wcoeffs")


def create_element(family, celltype, degree, lagrange_variant=LagrangeVariant.unset, dpc_variant=DPCVariant.unset,
                   discontinuous=False, dof_ordering=None, dtype=np.float64):
    """``basix.create_element`` (python/basix/finite_element.py:595-634)."""
    e = FiniteElement(_create_element(family, celltype, degree, lagrange_variant, dpc_variant, discontinuous,
                                      dof_ordering if dof_ordering is not None else [], np.dtype(dtype).char), dtype)
    if dof_ordering is not None and len(dof_ordering) and int(family) == ElementFamily.P \
            and int(celltype) in (CellType.quadrilateral, CellType.hexahedron):
        # tp_factors (finite-element.cpp:351-388): the ordering is the tensor-product one -> the factors are known
        if list(dof_ordering) == tp_dof_ordering(family, celltype, degree, lagrange_variant, dpc_variant, discontinuous):
            sub = create_element(ElementFamily.P, CellType.interval, degree, lagrange_variant, dpc_variant, True, None, dtype)
            e._tp_factors = [[sub] * (2 if int(celltype) == CellType.quadrilateral else 3)]
    return e


def create_custom_element(*args, **kwargs):
    raise NotBuiltNatively("create_custom_element: describe a custom element to the B200 library through "
                           "basix_b200.FiniteElement(...) / bx_element_create_from_arrays")


def tp_dof_ordering(family, celltype, degree, lagrange_variant=LagrangeVariant.unset, dpc_variant=DPCVariant.unset,
                    discontinuous=False):
    """``basix.tp_dof_ordering`` (finite-element.cpp:397-505): the dof numbering in which a P element on a quadrilateral
    / hexahedron is the tensor product of the (discontinuous) interval element with itself.  Derived here from the dof
    points instead of index formulas: dof (a, b[, c]) is the one sitting at (x[a], x[b][, x[c]]), x = the interval
    element's points in its own dof order, first index slowest."""
    if int(family) != ElementFamily.P or int(celltype) not in (CellType.quadrilateral, CellType.hexahedron):
        return None
    if degree == 0:
        return [0]
    sub = create_element(ElementFamily.P, CellType.interval, degree, lagrange_variant, dpc_variant, True)
    full = create_element(family, celltype, degree, lagrange_variant, dpc_variant, discontinuous)
    x1 = sub.points[:, 0]
    pts = full.points
    tdim = pts.shape[1]
    perm = []
    for idx in np.ndindex(*([len(x1)] * tdim)):
        target = np.array([x1[i] for i in idx])
        d = np.max(np.abs(pts - target[None]), axis=1)
        k = int(np.argmin(d))
        if d[k] > 1e-12:
            return None
        perm.append(k)
    if sorted(perm) != list(range(len(pts))):
        return None
    order = [0] * len(perm)
    for i, q in enumerate(perm):
        order[q] = i
    return order


def lex_dof_ordering(family, celltype, degree, lagrange_variant=LagrangeVariant.unset, dpc_variant=DPCVariant.unset,
                     discontinuous=False):
    """``basix.finite_element.lex_dof_ordering`` (finite-element.cpp:509-700): the numbering of a P element's dofs in
    lexicographic order of their points (last coordinate slowest).  Derived from the dof points themselves instead of the
    reference's index formulas, like ``tp_dof_ordering`` above: ``ordering[dof] = position of the dof's point``."""
    if int(family) != ElementFamily.P:
        raise RuntimeError("Lexicographic ordering is built for P elements only")
    e = create_element(family, celltype, degree,
                       LagrangeVariant.equispaced if int(lagrange_variant) == LagrangeVariant.unset and degree > 0
                       else lagrange_variant, dpc_variant, discontinuous)
    pts = np.round(np.asarray(e.points) * (4 * max(degree, 1)) * 64).astype(np.int64)  # exact keys for lattice points
    order = np.lexsort(tuple(pts[:, k] for k in range(pts.shape[1])))  # last key = last coordinate = slowest
    out = [0] * len(order)
    for position, dof in enumerate(order):
        out[int(dof)] = position
    return out


def create_tp_element(family, celltype, degree, lagrange_variant=LagrangeVariant.unset, dpc_variant=DPCVariant.unset,
                      discontinuous=False, dtype=np.float64):
    """``basix.create_tp_element`` (finite-element.cpp:327-341)."""
    order = tp_dof_ordering(family, celltype, degree, lagrange_variant, dpc_variant, discontinuous)
    if order is None:
        raise RuntimeError("Element does not have tensor product factorisation.")
    return create_element(family, celltype, degree, lagrange_variant, dpc_variant, discontinuous, order, dtype)
