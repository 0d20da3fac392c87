"""The C-ABI library loads and exports every symbol include/basix_b200.h declares; host-only entry points
behave (no compute calls: there is no GPU here, and no CPU fallback to fall into)."""

import ctypes as C
import os
import re

import numpy as np
import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "basix_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(bx_[a-zA-Z0-9_]+)\s*\(", text)))


def test_every_declared_symbol_is_exported_and_bound():
    from basix_b200 import _capi

    lib = C.CDLL(_capi.LIB_PATH)
    names = declared_symbols()
    assert len(names) >= 25
    for name in names:
        assert hasattr(lib, name), f"{name} declared in include/basix_b200.h but not exported"
    assert set(names) == set(_capi.SIGNATURES), "ctypes signatures out of sync with the header"


def test_host_only_entry_points():
    import basix_b200 as bx
    from basix_b200 import _capi

    lib = _capi.lib()
    assert lib.bx_version().decode().startswith("basix_b200")
    assert bx.polyset_dim(3, 0, 5) == 56 and bx.polyset_dim(5, 0, 4) == 125 and bx.polyset_nderivs(3, 2) == 10
    assert lib.bx_map_value_size(2, 0, 3, 3, 3) == 3 and lib.bx_map_value_size(4, 0, 4, 3, 2) == 9
    assert bx.index(1, 2) == 8 and bx.index(1, 1, 1) == 14


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device every compute call must fail loudly (BX_ERR_CUDA), never compute on the host."""
    import basix_b200 as bx

    if bx.device_count() > 0:
        pytest.skip("a CUDA device is present")
    e = bx.FiniteElement.from_arrays(np.load(os.path.join(ROOT, "tests", "golden", "elements", "P1_triangle_gll.npz")))
    assert e.dim == 3 and e.tabulate_shape(1, 5) == (3, 5, 3, 1)
    with pytest.raises(RuntimeError):
        e.tabulate(1, np.zeros((5, 2)))
    with pytest.raises(RuntimeError):
        bx.tabulate_polynomial_set(2, 0, 2, 1, np.zeros((5, 2)))
    e5 = bx.FiniteElement.from_arrays(np.load(os.path.join(ROOT, "tests", "golden", "elements", "P5_tetrahedron_gll.npz")))
    with pytest.raises(RuntimeError):
        e5.permute_batch(np.zeros((2, 56), dtype=np.int32), np.ones(2, dtype=np.uint32))
    # the entry points added for the callers either side of the path fail the same way
    with pytest.raises(RuntimeError):
        e5.permute_subentity_closure(np.arange(21, dtype=np.int32), 3, 2)
    en = bx.FiniteElement.from_arrays(np.load(os.path.join(ROOT, "tests", "golden", "elements", "N1curl3_tetrahedron_legendre.npz")))
    with pytest.raises(RuntimeError):
        en.interpolate_batch(np.zeros((2, en.interpolation_matrix.shape[1])))
    J = np.tile(np.eye(3), (2, 1, 1))
    with pytest.raises(RuntimeError):
        en.push_forward_broadcast(np.zeros((4, 3)), J, np.ones(2), J)
    # ... while the host tables stay inspectable
    assert e5.subentity_closure_size(2) == 21 and len(e5.subentity_closure_map(2, 3)) == 21
    # wrong point dimension is an argument error with the reference's message, device or not
    with pytest.raises(RuntimeError, match=r"Point dim \(3\) does not match element dim \(2\)\."):
        e.tabulate(1, np.zeros((5, 3)))


def test_product_package_does_not_import_the_oracle():
    import subprocess
    import sys

    code = ("import sys; sys.path.insert(0, %r); import basix_b200, basix_b200.sharding; "
            "assert not any(m == 'oracle' or m.startswith('oracle.') for m in sys.modules), 'oracle imported'" % ROOT)
    subprocess.run([sys.executable, "-c", code], check=True)
    for dirpath, _, files in os.walk(os.path.join(ROOT, "basix_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src, f
