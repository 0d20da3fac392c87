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


RT, N1E, LEG = 2, 3, 11
VECTOR_CASES = [(RT, TRI, 1, LEG), (RT, TRI, 2, LEG), (RT, TRI, 3, EQ), (RT, TET, 1, LEG), (RT, TET, 2, LEG),
                (RT, TET, 2, 0), (RT, TET, 3, GLL), (RT, TET, 4, LEG), (N1E, TRI, 1, LEG), (N1E, TRI, 2, LEG),
                (N1E, TRI, 3, LEG), (N1E, TET, 1, LEG), (N1E, TET, 2, GLL), (N1E, TET, 3, LEG), (N1E, TET, 4, LEG)]


@pytest.mark.parametrize("family,cell,degree,variant", VECTOR_CASES)
def test_rt_and_nedelec_match_reference(ref, family, cell, degree, variant):
    """Native RT / N1curl (quadrature, moment dofs, orthogonalised spaces, Piola-mapped entity transformations)
    against the reference: coefficient matrix, entity dofs, transformation matrices, and the composed net operator
    of T_apply for random cell_info."""
    import basix_b200 as bx

    r = ref.RefElement(family, cell, degree, variant)
    e = bx.create_element(family, cell, degree, variant)
    assert e.dim == r.dim and e.value_size == r.value_size and int(e.map_type) == r.map_type
    assert e.entity_dofs == r.entity_dofs
    assert e.dof_transformations_are_permutations == r.dof_transformations_are_permutations
    C, Cr = e.coefficient_matrix, r.coefficient_matrix
    assert np.max(np.abs(C - Cr)) <= 1e-12 * np.max(np.abs(Cr))
    names = {1: "interval", 2: "triangle", 4: "quadrilateral"}
    et = e.entity_transformations()
    assert sorted(et) == sorted(names[k] for k in r.entity_transformations)
    for k, m in r.entity_transformations.items():
        assert et[names[k]].shape == m.shape
        assert np.max(np.abs(et[names[k]] - m), initial=0.0) <= 1e-12
    rng = np.random.default_rng(degree)
    for ci in rng.integers(0, 2 ** 30, 4):
        u = rng.standard_normal((1, r.dim))
        want = r.T_apply_batch("T_apply", u.copy(), 1, np.array([ci], dtype=np.uint32))[0]
        got = e.net_operator(int(ci)) @ u[0]
        assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))


def test_vector_families_errors_and_discontinuous(ref):
    import basix_b200 as bx

    with pytest.raises(RuntimeError, match="need to be given a variant"):
        bx.create_element(bx.ElementFamily.RT, TET, 4)  # moment space of degree 3 needs a variant
    with pytest.raises(RuntimeError):
        bx.create_element(bx.ElementFamily.RT, QUAD, 1, LEG)
    with pytest.raises(RuntimeError, match="built natively"):
        bx.create_element(bx.ElementFamily.BDM, TRI, 1, LEG)
    e = bx.create_element(bx.ElementFamily.N1E, TET, 2, LEG, discontinuous=True)
    r = ref.RefElement(N1E, TET, 2, LEG, 0, True)
    assert e.entity_dofs == r.entity_dofs and e.dof_transformations_are_identity
    assert np.max(np.abs(e.coefficient_matrix - r.coefficient_matrix)) <= 1e-12 * np.max(np.abs(r.coefficient_matrix))


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
