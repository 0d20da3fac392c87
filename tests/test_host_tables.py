"""Host-side DOF-transformation tables (composed per cell_info bit-field) vs the reference's
sequential application -- no GPU needed.  Reference: finite-element.h:1703-1998."""

import numpy as np
import pytest

from conftest import make_pair

P, RT, N1E = 1, 2, 3
INTERVAL, TRI, TET, QUAD, HEX = 1, 2, 3, 4, 5
GLL, LEG = 2, 11

OPS = ["T_apply", "Tt_apply", "Tt_inv_apply", "Tinv_apply"]


def cell_infos(n=24, seed=3):
    rng = np.random.default_rng(seed)
    fixed = [0, 1, 2, 4, 6, 7, 0x1000, 0x2ABCD, 0x3FFFFFFF, 0xFFFFFFFF]
    return np.array(fixed + list(rng.integers(0, 2 ** 30, n)), dtype=np.uint32)


@pytest.mark.parametrize("cell,degree", [(TRI, 1), (TRI, 3), (TRI, 4), (TET, 3), (TET, 4), (TET, 5), (QUAD, 3),
                                         (HEX, 2), (HEX, 3), (HEX, 4), (INTERVAL, 3)])
def test_permutation_tables_match_reference(ref, cell, degree):
    r, e = make_pair(ref, P, cell, degree, GLL)
    assert e.dof_transformations_are_permutations == r.dof_transformations_are_permutations
    assert e.dof_transformations_are_identity == r.dof_transformations_are_identity
    ci = cell_infos()
    base = np.tile(np.arange(r.dim, dtype=np.int32), (len(ci), 1))
    for inverse, op in ((False, "T_apply"), (True, "Tt_apply")):
        want = r.permute_batch(base.copy(), ci, inverse=inverse)
        for k, c in enumerate(ci):
            assert np.array_equal(e.net_permutation(int(c), op), want[k]), (op, hex(int(c)))
    # all four operator kinds through T_apply on int32 data
    for op in OPS:
        want = r.T_apply_batch(op, base.copy(), 1, ci)
        for k, c in enumerate(ci):
            assert np.array_equal(e.net_permutation(int(c), op), want[k]), (op, hex(int(c)))


@pytest.mark.parametrize("family,cell,degree", [(RT, TET, 2), (RT, TET, 4), (N1E, TET, 3), (N1E, TET, 2),
                                                (RT, TRI, 3), (N1E, TRI, 2), (N1E, HEX, 2), (RT, HEX, 2)])
def test_matrix_tables_match_reference(ref, family, cell, degree):
    r, e = make_pair(ref, family, cell, degree, LEG)
    assert e.dof_transformations_are_permutations == r.dof_transformations_are_permutations
    assert e.dof_transformations_are_identity == r.dof_transformations_are_identity
    n = r.dim
    ci = cell_infos(8)
    for op in OPS:
        # columns of the identity through the reference = the dense net operator
        for c in ci:
            eye = np.ascontiguousarray(np.eye(n).T.reshape(1, -1))  # data[dof*n + b], block size n
            got_ref = r.T_apply_batch(op, eye.copy(), n, np.array([c], dtype=np.uint32)).reshape(n, n)
            A = e.net_operator(int(c), op)
            assert np.max(np.abs(A - got_ref)) <= 1e-13 * max(1.0, np.max(np.abs(got_ref))), (op, hex(int(c)))


def test_right_variants_share_operators(ref):
    """X_right acts on the transposed layout with the operator of its left twin
    (finite-element.h:1904-1998)."""
    r, e = make_pair(ref, N1E, TET, 2, LEG)
    n = r.dim
    twin = {"Tt_apply_right": "T_apply", "Tinv_apply_right": "Tt_inv_apply", "T_apply_right": "Tt_apply",
            "Tt_inv_apply_right": "Tinv_apply"}
    rng = np.random.default_rng(0)
    for c in cell_infos(4):
        for right, left in twin.items():
            bs = 3
            data = rng.standard_normal((bs, n))  # right layout: bs contiguous blocks of ndofs
            want = r.T_apply_batch(right, np.ascontiguousarray(data.reshape(1, -1)).copy(), bs,
                                   np.array([c], dtype=np.uint32)).reshape(bs, n)
            A = e.net_operator(int(c), right)
            assert np.allclose(want, data @ A.T, rtol=0, atol=1e-13)
            assert np.array_equal(A, e.net_operator(int(c), left))


@pytest.mark.parametrize("family,cell,degree,variant", [(P, TRI, 3, GLL), (P, TET, 4, GLL), (P, QUAD, 2, GLL), (P, HEX, 3, GLL),
                                                        (RT, TET, 2, LEG), (N1E, TET, 3, LEG), (N1E, TRI, 2, LEG), (P, 6, 2, 0),
                                                        (P, 7, 2, 0), (P, INTERVAL, 4, GLL)])
def test_metadata_matches_reference(ref, family, cell, degree, variant):
    """num_entity_dofs, entity_closure_dofs, base_transformations, interpolation_is_identity, == (host only)."""
    r, e = make_pair(ref, family, cell, degree, variant)
    assert e.num_entity_dofs == [[len(v) for v in row] for row in r.entity_dofs]
    assert e.entity_closure_dofs == r.entity_closure_dofs
    assert e.num_entity_closure_dofs == [[len(v) for v in row] for row in r.entity_closure_dofs]
    bt, want = e.base_transformations(), r.base_transformations()
    assert bt.shape == want.shape and np.array_equal(bt, want)
    assert e.interpolation_is_identity == (family == P)
    assert e.dtype is np.float64 and e == e and hash(e) == hash(e)
    r2, e2 = make_pair(ref, P, TRI, 1)
    assert not (e == e2)


def test_tabulate_dispatch_of_the_baseline_configs():
    """Which kernel tabulate() takes is decided on the host (no device needed): BASELINE configs and fall-backs."""
    import basix_b200 as bx

    assert bx.create_element(P, TRI, 1, GLL).tabulate_path(1) == "register"                 # cfg 1
    p5 = bx.create_element(P, TET, 5, GLL)
    assert p5.tabulate_path(2) == "derivative-operator"                                     # cfg 2
    assert bx.create_element(P, HEX, 4, GLL).tabulate_path(1) == "tensor-product"           # cfg 3
    assert bx.create_element(N1E, TET, 3, LEG).tabulate_path(1) == "derivative-operator"    # vector valued
    p6 = bx.create_element(P, TET, 6, GLL)
    assert p6.tabulate_path(0) == "derivative-operator" and p6.tabulate_path(1) == "fused"  # operators too large
    with bx.tabulate_paths("fused"):
        assert p5.tabulate_path(2) == "fused"
    with bx.tabulate_paths():
        assert p5.tabulate_path(2) == "generic"
    assert p5.tabulate_path(2) == "derivative-operator"
