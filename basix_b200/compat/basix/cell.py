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
