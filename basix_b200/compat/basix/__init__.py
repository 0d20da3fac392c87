"""Drop-in ``basix`` package backed by ``basix_b200`` (CUDA, sm_100a).

Mirrors the names exported by python/basix/__init__.py for the hot path of the reference -- ``create_element``,
``FiniteElement.tabulate / push_forward / pull_back / T_apply ...`` -- plus the small host helpers its callers and tests
use next to them (``topology``, ``geometry``, ``index``, ``create_lattice``, ``compute_interpolation_operator``).
Everything else raises :class:`NotBuiltNatively`.
"""

from basix_b200.finite_element import DPCVariant, LagrangeVariant, MapType, PolysetType

from basix import _basixcpp, cell, finite_element, interpolation, lattice, polynomials, quadrature, sobolev_spaces  # noqa: F401
from basix._basixcpp import NotBuiltNatively
from basix.cell import CellType, geometry, topology
from basix.finite_element import ElementFamily, create_custom_element, create_element, create_tp_element
from basix.interpolation import compute_interpolation_operator
from basix.lattice import LatticeSimplexMethod, LatticeType, create_lattice
from basix.polynomials import PolynomialType, tabulate_polynomials
from basix.sobolev_spaces import SobolevSpace
from basix.utils import index

__version__ = "0.12.0.dev0+b200"


from basix.quadrature import QuadratureType, make_quadrature  # noqa: E402

__all__ = [
    "CellType", "DPCVariant", "ElementFamily", "LagrangeVariant", "LatticeSimplexMethod", "LatticeType", "MapType",
    "NotBuiltNatively", "PolynomialType", "PolysetType", "QuadratureType", "SobolevSpace", "__version__", "cell",
    "compute_interpolation_operator", "create_custom_element", "create_element", "create_lattice", "create_tp_element",
    "finite_element", "geometry", "index", "lattice", "make_quadrature", "polynomials", "sobolev_spaces",
    "tabulate_polynomials", "topology",
]
