"""Native create_element (csrc/create.cu, host set-up code) against the compiled reference: coefficient
matrices, entity dofs and entity transformations of the Lagrange family.  No GPU needed: element objects build
their host tables without a device."""

import numpy as np
import pytest

P = 1
INTERVAL, TRI, TET, QUAD, HEX = 1, 2, 3, 4, 5
EQ, GLL = 1, 2

CASES = [(c, d, v) for c in (INTERVAL, TRI, TET, QUAD, HEX) for d in (1, 2, 3, 4, 5) for v in (EQ, GLL)
         if not (c == HEX and d == 5)] + [(TET, 6, GLL), (TRI, 7, GLL), (INTERVAL, 9, GLL), (QUAD, 6, GLL)]


@pytest.mark.parametrize("cell,degree,variant", CASES)
def test_lagrange_matches_reference(ref, cell, degree, variant):
    import basix_b200 as bx

    r = ref.RefElement(P, cell, degree, variant)
    e = bx.create_element(bx.ElementFamily.P, cell, degree, variant)
    assert e.dim == r.dim and e.value_size == 1 and int(e.map_type) == r.map_type
    assert e.entity_dofs == r.entity_dofs
    assert e.dof_transformations_are_permutations == r.dof_transformations_are_permutations
    assert e.dof_transformations_are_identity == r.dof_transformations_are_identity
    C, Cr = e.coefficient_matrix, r.coefficient_matrix
    assert C.shape == Cr.shape
    assert np.max(np.abs(C - Cr)) <= 1e-11 * np.max(np.abs(Cr))
    names = {1: "interval", 2: "triangle", 4: "quadrilateral"}
    et = e.entity_transformations()
    assert sorted(et) == sorted(names[k] for k in r.entity_transformations)
    for k, m in r.entity_transformations.items():
        assert et[names[k]].shape == m.shape
        assert np.max(np.abs(et[names[k]] - m), initial=0.0) <= 1e-11
    # the composed permutation tables agree for a few cell_info values
    if r.dof_transformations_are_permutations and not r.dof_transformations_are_identity:
        for ci in (0x2ABCD, 0x15F3A9, 0x3FFFFFFF):
            d = np.arange(r.dim, dtype=np.int32).reshape(1, -1)
            want = r.permute_batch(d.copy(), np.array([ci], dtype=np.uint32))[0]
            src = e.net_permutation(ci)
            assert np.array_equal(np.arange(r.dim, dtype=np.int32)[src], want)


def test_variants_discontinuous_and_ordering(ref):
    import basix_b200 as bx

    # unset resolves to gll_warped below degree 3, and is an error above (e-lagrange.cpp:491-500)
    e = bx.create_element(bx.ElementFamily.P, TRI, 2)
    r = ref.RefElement(P, TRI, 2, 0)
    assert np.allclose(e.coefficient_matrix, r.coefficient_matrix, rtol=0, atol=1e-12)
    with pytest.raises(RuntimeError, match="need to be given a variant"):
        bx.create_element(bx.ElementFamily.P, TRI, 3)
    with pytest.raises(RuntimeError, match="continuous order 0"):
        bx.create_element(bx.ElementFamily.P, TRI, 0, EQ)
    with pytest.raises(RuntimeError, match="only the P"):
        bx.create_element(bx.ElementFamily.RT, TRI, 1)
    # discontinuous: all dofs interior, no transformations left
    e = bx.create_element(bx.ElementFamily.P, TET, 3, GLL, discontinuous=True)
    r = ref.RefElement(P, TET, 3, GLL, 0, True)
    assert e.entity_dofs == r.entity_dofs and e.dof_transformations_are_identity
    assert np.max(np.abs(e.coefficient_matrix - r.coefficient_matrix)) <= 1e-11 * np.max(np.abs(r.coefficient_matrix))
    e0 = bx.create_element(bx.ElementFamily.P, HEX, 0, GLL, discontinuous=True)
    r0 = ref.RefElement(P, HEX, 0, GLL, 0, True)
    assert e0.dim == 1 and np.allclose(e0.coefficient_matrix, r0.coefficient_matrix)
    # dof_ordering
    order = [int(v) for v in np.random.default_rng(0).permutation(10)]
    e = bx.create_element(bx.ElementFamily.P, TRI, 3, GLL, dof_ordering=order)
    r = ref.RefElement(P, TRI, 3, GLL, 0, False, order)
    assert e.entity_dofs == r.entity_dofs and e.dof_ordering == [int(v) for v in r.dof_ordering]
    with pytest.raises(RuntimeError, match="Incorrect number of dofs"):
        bx.create_element(bx.ElementFamily.P, TRI, 3, GLL, dof_ordering=[0, 1])


def test_native_element_is_a_partition_of_unity_basis():
    """Without the reference: C P(x_dofs) = I means sum of coefficients against the constant mode is 1."""
    import basix_b200 as bx

    e = bx.create_element(bx.ElementFamily.P, TET, 4, GLL)
    C = e.coefficient_matrix
    # the first orthonormal function on the tetrahedron is the constant sqrt(6): sum_i phi_i = 1
    assert abs(C[:, 0].sum() * np.sqrt(6.0) - 1.0) < 1e-12
    assert np.max(np.abs(C[:, 1:].sum(axis=0))) < 1e-11
