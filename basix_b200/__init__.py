"""basix_b200 -- B200-native (sm_100a) implementation of the FEniCS/basix hot path:
``FiniteElement.tabulate``, ``push_forward`` / ``pull_back`` and the per-cell DOF
transformations, behind the reference's Python names.

Everything is computed by hand-written CUDA kernels in ``libbasix_b200.so`` (C ABI in
``include/basix_b200.h``).  There is no CPU fallback: without the built library, or without a
CUDA device, the compute calls raise ``RuntimeError``.
"""

from . import _capi
from .finite_element import (
    CellType,
    DPCVariant,
    ElementFamily,
    FiniteElement,
    LagrangeVariant,
    MapType,
    PolysetType,
    compute_interpolation_operator,
    create_element,
    geometry,
    polyset_dim,
    polyset_nderivs,
    tabulate_polynomial_set,
)

__version__ = "0.1.0"


def device_count() -> int:
    return _capi.lib().bx_device_count()


def set_device(device: int) -> None:
    _capi.check(_capi.lib().bx_set_device(int(device)))


def kernel_launch_count() -> int:
    return int(_capi.lib().bx_kernel_launch_count())


class tabulate_paths:
    """Context manager narrowing which tabulate kernels may be used (tests / profiling): names from
    ``FiniteElement.tabulate_path`` -- "fused", "tensor-product", "register", "derivative-operator"; the generic
    path is always available.  "tensor-product-staged" additionally allows the staged variant of the
    tensor-product kernel (same path name; on by default)."""

    _BITS = {"generic": 0, "fused": 1, "tensor-product": 2, "register": 3, "derivative-operator": 4,
             "tensor-product-staged": 5}

    def __init__(self, *allowed: str):
        self.mask = 1
        for name in allowed:
            self.mask |= 1 << self._BITS[name]

    def __enter__(self):
        self.prev = _capi.lib().bx_tabulate_path_mask(self.mask)
        return self

    def __exit__(self, *exc):
        _capi.lib().bx_tabulate_path_mask(self.prev)
        return False


def index(p: int, q: int | None = None, r: int | None = None) -> int:
    """``basix.index`` (python/basix/utils.py; cpp/basix/indexing.h:13-34)."""
    if q is None:
        return p
    if r is None:
        return (p + q + 1) * (p + q) // 2 + q
    return (p + q + r) * (p + q + r + 1) * (p + q + r + 2) // 6 + (q + r) * (q + r + 1) // 2 + r


__all__ = [
    "CellType", "DPCVariant", "compute_interpolation_operator", "create_element", "ElementFamily", "FiniteElement", "LagrangeVariant", "MapType", "PolysetType",
    "device_count", "geometry", "index", "kernel_launch_count", "polyset_dim", "polyset_nderivs", "set_device", "tabulate_paths",
    "tabulate_polynomial_set",
]
