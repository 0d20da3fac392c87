"""``basix.cell`` (python/basix/cell.py): topology / geometry of the reference cells."""

from basix._basixcpp import CellType, sub_entity_connectivity, sub_entity_type  # noqa: F401
from basix._basixcpp import cell_edge_jacobians as edge_jacobians  # noqa: F401
from basix._basixcpp import cell_facet_jacobians as facet_jacobians  # noqa: F401
from basix._basixcpp import geometry as _geometry
from basix._basixcpp import topology as _topology


def topology(celltype):
    """Vertex numbers of every sub-entity, per dimension (cell::topology, cell.cpp:52-141)."""
    return _topology(celltype)


def geometry(celltype):
    """Vertex coordinates of the reference cell (cell::geometry, cell.cpp:20-50)."""
    return _geometry(celltype)


def subentity_types(celltype):
    return [[sub_entity_type(celltype, d, i) for i in range(len(row))] for d, row in enumerate(_topology(celltype))]


# ---- host helpers of python/basix/cell.py built from topology / geometry (cell.cpp:404-596) ----------------------
import numpy as _np

_VOLUME = {0: 0.0, 1: 1.0, 2: 0.5, 3: 1.0 / 6.0, 4: 1.0, 5: 1.0, 6: 0.5, 7: 1.0 / 3.0}


def volume(celltype) -> float:
    """Volume of the reference cell (cell::volume, cell.cpp:404-427)."""
    return _VOLUME[int(celltype)]


def facet_normals(celltype):
    """Unit normals of the facets, oriented by the facet's vertex order (cell::facet_normals, cell.cpp:450-510)."""
    x = _np.asarray(_geometry(celltype), dtype=_np.float64)
    tdim = x.shape[1]
    facets = _topology(celltype)[tdim - 1]
    n = _np.ones((len(facets), tdim))
    for f, facet in enumerate(facets):
        if tdim == 2:
            n[f] = [x[facet[1], 1] - x[facet[0], 1], x[facet[0], 0] - x[facet[1], 0]]
        elif tdim == 3:
            n[f] = _np.cross(x[facet[1]] - x[facet[0]], x[facet[2]] - x[facet[0]])
        if tdim > 1:
            n[f] /= _np.sqrt(_np.dot(n[f], n[f]))
    return n


def facet_orientations(celltype):
    """True where ``facet_normals`` points into the cell (cell::facet_orientations, cell.cpp:556-584)."""
    x = _np.asarray(_geometry(celltype), dtype=_np.float64)
    facets = _topology(celltype)[x.shape[1] - 1]
    midpoint = x.sum(axis=0) / x.shape[0]
    return [bool(_np.dot(n, x[facet[0]] - midpoint) < 0) for n, facet in zip(facet_normals(celltype), facets)]


def facet_outward_normals(celltype):
    """Unit normals pointing out of the cell (cell::facet_outward_normals, cell.cpp:431-446)."""
    n = facet_normals(celltype)
    for f, flip in enumerate(facet_orientations(celltype)):
        if flip:
            n[f] = -n[f]
    return n


def facet_reference_volumes(celltype):
    """Volumes of the reference cells of the facets (cell::facet_reference_volumes, cell.cpp:587-596)."""
    tdim = _np.asarray(_geometry(celltype)).shape[1]
    return _np.array([volume(t) for t in subentity_types(celltype)[tdim - 1]])
