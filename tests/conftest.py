"""Shared test plumbing.

* ``-m "not gpu"``: oracle vs golden vectors, host tables, C-ABI symbol export (no compute calls).
* ``-m gpu``: parity tests proper -- CUDA path (through the C ABI) vs the oracle.

The oracle (``oracle/``) is the checker only; the product package never imports it.
"""

import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def pytest_collection_modifyitems(config, items):
    # a gpu-marked test on a box without a GPU is a collection error elsewhere; here we skip with a reason
    import basix_b200

    if basix_b200.device_count() > 0:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


@pytest.fixture(scope="session")
def ref():
    from oracle import ref as _ref

    if not _ref.available():
        pytest.skip("reference oracle library not built (oracle/build_ref.sh)")
    return _ref


_ELEMENT_CACHE = {}


def make_pair(ref, family, cell, degree, lvariant=0, dvariant=0, discontinuous=False, dof_ordering=None):
    """(reference element, basix_b200 element built from the reference element's arrays)."""
    import basix_b200

    key = (family, cell, degree, lvariant, dvariant, discontinuous, tuple(dof_ordering or ()))
    if key not in _ELEMENT_CACHE:
        r = ref.RefElement(family, cell, degree, lvariant, dvariant, discontinuous, dof_ordering)
        e = basix_b200.FiniteElement(
            cell_type=r.cell_type, polyset_type=r.polyset_type, degree=r.degree,
            embedded_superdegree=r.embedded_superdegree, map_type=r.map_type, value_shape=r.value_shape,
            coefficient_matrix=r.coefficient_matrix, entity_dofs=r.entity_dofs,
            entity_transformations=r.entity_transformations, dof_ordering=r.dof_ordering, family=family,
            lagrange_variant=lvariant, dpc_variant=dvariant, discontinuous=discontinuous, points=r.points,
            interpolation_matrix=r.interpolation_matrix)
        _ELEMENT_CACHE[key] = (r, e)
    return _ELEMENT_CACHE[key]


def random_points(cell, n, seed=42):
    """Uniform points in the reference cell."""
    rng = np.random.default_rng(seed)
    tdim = {1: 1, 2: 2, 3: 3, 4: 2, 5: 3, 6: 3, 7: 3}[cell]
    u = rng.random((n, tdim))
    if cell == 6:  # prism: triangle x interval
        flip = u[:, :2].sum(axis=1) > 1
        u[flip, :2] = 1 - u[flip, :2]
    elif cell == 7:  # pyramid: (x, y) in [0, 1 - z]^2
        u[:, 2] = 1 - np.cbrt(u[:, 2])
        u[:, :2] *= (1 - u[:, 2])[:, None]
    if cell == 2:
        flip = u.sum(axis=1) > 1
        u[flip] = 1 - u[flip]
    elif cell == 3:
        # sort-and-difference: uniform in the simplex
        s = np.sort(rng.random((n, 3)), axis=1)
        u = np.stack([s[:, 0], s[:, 1] - s[:, 0], s[:, 2] - s[:, 1]], axis=1)
    return np.ascontiguousarray(u)


def slab_close(a, b, rtol=1e-12):
    """Parity metric of SURVEY 7.2-4: max|a-b| <= rtol * max|b| per leading-index slab."""
    a, b = np.asarray(a), np.asarray(b)
    assert a.shape == b.shape, (a.shape, b.shape)
    for d in range(a.shape[0]) if a.ndim > 1 else [slice(None)]:
        scale = np.max(np.abs(b[d])) if b[d].size else 0.0
        err = np.max(np.abs(a[d] - b[d])) if b[d].size else 0.0
        assert err <= rtol * max(scale, 1e-300), f"slab {d}: err {err:.3e} > {rtol:g} * {scale:.3e}"
