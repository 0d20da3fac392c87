"""ctypes binding of the C ABI declared in ``include/basix_b200.h``.

The shared library is built in-tree (``basix_b200/libbasix_b200.so``) by
``basix_b200/csrc/Makefile`` and is the ONLY compute backend: importing this
module never falls back to numpy/torch arithmetic.  On a machine without a GPU
the library still loads (static cudart) so symbols can be inspected, but every
compute entry point returns ``BX_ERR_CUDA``.
"""

from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libbasix_b200.so")

BX_MEM_HOST = 0
BX_MEM_DEVICE = 1


class ElementDesc(C.Structure):
    """``bx_element_desc`` (include/basix_b200.h)."""

    _fields_ = [
        ("cell_type", C.c_int),
        ("polyset_type", C.c_int),
        ("degree", C.c_int),
        ("embedded_superdegree", C.c_int),
        ("map_type", C.c_int),
        ("ndofs", C.c_int),
        ("value_size", C.c_int),
        ("coeffs", C.c_void_p),
        ("dof_ordering", C.c_void_p),
        ("n_dof_ordering", C.c_int),
        ("entity_dof_offsets", C.c_void_p),
        ("entity_dofs", C.c_void_p),
        ("etrans_interval", C.c_void_p),
        ("etrans_interval_dim", C.c_int),
        ("etrans_triangle", C.c_void_p),
        ("etrans_triangle_dim", C.c_int),
        ("etrans_quadrilateral", C.c_void_p),
        ("etrans_quadrilateral_dim", C.c_int),
        ("points", C.c_void_p),
        ("npts", C.c_int),
        ("interpolation_matrix", C.c_void_p),
    ]


# name -> (restype, argtypes); every symbol include/basix_b200.h declares
_SIZE = C.c_size_t
_P = C.c_void_p
_I = C.c_int
SIGNATURES = {
    "bx_last_error": (C.c_char_p, []),
    "bx_version": (C.c_char_p, []),
    "bx_device_count": (_I, []),
    "bx_set_device": (_I, [_I]),
    "bx_stream_synchronize": (_I, [_P]),
    "bx_kernel_launch_count": (C.c_uint64, []),
    "bx_polyset_dim": (_I, [_I, _I, _I]),
    "bx_polyset_nderivs": (_I, [_I, _I]),
    "bx_polyset_tabulate_f64": (_I, [_I, _I, _I, _I, _P, _SIZE, _P, _I, _P]),
    "bx_polyset_tabulate_f32": (_I, [_I, _I, _I, _I, _P, _SIZE, _P, _I, _P]),
    "bx_element_create_from_arrays": (_I, [C.POINTER(ElementDesc), C.POINTER(_P)]),
    "bx_element_destroy": (None, [_P]),
    "bx_create_element": (_I, [_I, _I, _I, _I, _I, _I, _P, _I, C.POINTER(_P)]),
    "bx_element_coefficient_matrix": (_I, [_P, _P]),
    "bx_element_entity_dofs": (_I, [_P, _I, _I, _P]),
    "bx_element_entity_transformations": (_I, [_P, _I, _P, _P]),
    "bx_element_dof_ordering": (_I, [_P, _P]),
    "bx_element_info": (_I, [_P, _P]),
    "bx_element_tabulate_shape": (_I, [_P, _I, _SIZE, _P]),
    "bx_element_tabulate_path": (_I, [_P, _I]),
    "bx_element_prepare": (_I, [_P, _I]),
    "bx_tabulate_path_mask": (C.c_uint, [C.c_uint]),
    "bx_element_net_operator": (_I, [_P, _I, C.c_uint32, _P, _P]),
    "bx_tabulate_f64": (_I, [_P, _I, _P, _SIZE, _I, _P, _I, _P]),
    "bx_tabulate_f32": (_I, [_P, _I, _P, _SIZE, _I, _P, _I, _P]),
    "bx_map_f32": (_I, [_I, _I, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_push_forward_f32": (_I, [_P, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_push_forward_broadcast_f64": (_I, [_P, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_push_forward_broadcast_f32": (_I, [_P, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_pull_back_f32": (_I, [_P, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_map_value_size": (_I, [_I, _I, _I, _I, _I]),
    "bx_map_f64": (_I, [_I, _I, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_push_forward_f64": (_I, [_P, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_pull_back_f64": (_I, [_P, _P, _P, _P, _P, _SIZE, _SIZE, _I, _I, _I, _P, _I, _P]),
    "bx_geometry_f64": (_I, [_P, _P, _SIZE, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "bx_geometry_f32": (_I, [_P, _P, _SIZE, _I, _I, _I, _I, _P, _P, _P, _I, _P]),
    "bx_push_forward_fused_f64": (_I, [_P, _I, _P, _SIZE, _P, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_push_forward_fused_f32": (_I, [_P, _I, _P, _SIZE, _P, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_physical_basis_f64": (_I, [_P, _I, _P, _SIZE, _I, _P, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_physical_basis_pointwise_f64": (_I, [_P, _I, _P, _SIZE, _I, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_push_forward_fused_pointwise_f64": (_I, [_P, _I, _P, _SIZE, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_push_forward_fused_pointwise_f32": (_I, [_P, _I, _P, _SIZE, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_T_apply_f64": (_I, [_P, _I, _P, _SIZE, _SIZE, _I, _P, _I, _P]),
    "bx_T_apply_f32": (_I, [_P, _I, _P, _SIZE, _SIZE, _I, _P, _I, _P]),
    "bx_T_apply_i32": (_I, [_P, _I, _P, _SIZE, _SIZE, _I, _P, _I, _P]),
    "bx_permute_i32": (_I, [_P, _I, _P, _SIZE, _P, _I, _P]),
    "bx_element_interpolation_shape": (_I, [_P, _P]),
    "bx_element_points": (_I, [_P, _P]),
    "bx_element_interpolation_matrix": (_I, [_P, _P]),
    "bx_interpolate_f64": (_I, [_P, _I, _P, _SIZE, _P, _I, _P]),
    "bx_interpolate_f32": (_I, [_P, _I, _P, _SIZE, _P, _I, _P]),
    "bx_interpolate_pulled_f64": (_I, [_P, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_interpolate_pulled_f32": (_I, [_P, _P, _P, _P, _P, _SIZE, _I, _P, _I, _P]),
    "bx_compute_interpolation_operator_f64": (_I, [_P, _P, _P, _P, _I, _P]),
    "bx_make_quadrature": (_I, [_I, _I, _P, _P, C.POINTER(C.c_size_t)]),
    "bx_gauss_jacobi_rule": (_I, [C.c_double, _I, _P, _P]),
    "bx_create_lattice": (_I, [_I, _I, _I, _I, _I, _P, C.POINTER(C.c_size_t)]),
    "bx_permute_subentity_closure_i32": (_I, [_P, _I, _P, _SIZE, _SIZE, _P, _I, _I, _I, _P]),
    "bx_element_subentity_closure_size": (_I, [_P, _I]),
    "bx_element_subentity_closure_map": (_I, [_P, _I, _I, C.c_uint32, _P]),
}

_lib = None


def lib() -> C.CDLL:
    """Load the CUDA library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"basix_b200: CUDA library missing ({LIB_PATH}); build it with "
                "`python -c 'import __graft_entry__ as g; g.build()'` or `make -C basix_b200/csrc`. "
                "There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, (res, args) in SIGNATURES.items():
            fn = getattr(L, name)
            fn.restype = res
            fn.argtypes = args
        _lib = L
    return _lib


def check(rc: int) -> None:
    """Turn a non-zero status into RuntimeError carrying the library's message
    (the reference raises std::runtime_error -> RuntimeError)."""
    if rc != 0:
        msg = lib().bx_last_error().decode()
        raise RuntimeError(msg or f"basix_b200 error {rc}")
