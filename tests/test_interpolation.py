"""points() / interpolation_matrix() (reference finite-element.h:1167-1170, 1222-1226) and the batched product
``coefficients = interpolation_matrix @ values`` the reference leaves to the caller (finite-element.h:1183-1216).

CPU part: the native factory's points / matrix against the compiled reference.  The native RT / N1curl elements
use Gauss-Jacobi rules where the reference uses Xiao-Gimbutas tables, so their points differ while
the functionals agree: both matrices are checked to give the same coefficients for polynomial fields, and to be the
dual of the element's own basis.  GPU part: the kernel against numpy with the reference's matrix, 1e-12 relative.
"""

import numpy as np
import pytest

from conftest import make_pair

P, RT, N1E = 1, 2, 3
INTERVAL, TRI, TET, QUAD, HEX = 1, 2, 3, 4, 5
EQ, GLL, LEG = 1, 2, 11


@pytest.mark.parametrize("cell,degree,variant", [(INTERVAL, 4, GLL), (TRI, 1, 0), (TRI, 4, GLL), (TET, 5, GLL), (QUAD, 3, EQ),
                                                 (HEX, 2, 0), (HEX, 4, GLL)])
def test_native_lagrange_points_and_matrix(ref, cell, degree, variant):
    import basix_b200 as bx

    r = ref.RefElement(P, cell, degree, variant)
    e = bx.create_element(P, cell, degree, variant)
    assert e.points.shape == r.points.shape and np.max(np.abs(e.points - r.points)) <= 1e-13
    assert np.array_equal(e.interpolation_matrix, r.interpolation_matrix)  # the identity


def test_native_lagrange_points_follow_dof_ordering(ref):
    import basix_b200 as bx

    order = [int(v) for v in np.random.default_rng(1).permutation(10)]
    r = ref.RefElement(P, TRI, 3, GLL, 0, False, order)
    e = bx.create_element(P, TRI, 3, GLL, dof_ordering=order)
    assert np.max(np.abs(e.points - r.points)) <= 1e-13


def _poly_field(x, vs, degree, seed):
    """Random polynomial vector field of total degree <= degree at the points x: (npts, vs)."""
    rng = np.random.default_rng(seed)
    tdim = x.shape[1]
    out = np.zeros((x.shape[0], vs))
    for j in range(vs):
        for _ in range(6):
            pw = rng.integers(0, degree + 1, tdim)
            while pw.sum() > degree:
                pw = rng.integers(0, degree + 1, tdim)
            out[:, j] += rng.standard_normal() * np.prod(x ** pw, axis=1)
    return out


@pytest.mark.parametrize("family,cell,degree,variant", [(RT, TRI, 2, LEG), (RT, TET, 1, LEG), (RT, TET, 4, LEG), (N1E, TRI, 3, LEG),
                                                        (N1E, TET, 2, GLL), (N1E, TET, 3, LEG)])
def test_native_vector_interpolation_functionals(ref, family, cell, degree, variant):
    import basix_b200 as bx

    r = ref.RefElement(family, cell, degree, variant)
    e = bx.create_element(family, cell, degree, variant)
    vs, n = r.value_size, r.dim
    assert e.interpolation_matrix.shape == (n, vs * e.points.shape[0])
    for seed in range(3):
        fe = _poly_field(e.points, vs, degree, seed).T.reshape(-1)  # component major: values[j * npts + p]
        fr = _poly_field(r.points, vs, degree, seed).T.reshape(-1)
        want = r.interpolation_matrix @ fr
        got = e.interpolation_matrix @ fe
        assert np.max(np.abs(got - want)) <= 1e-11 * max(1.0, np.max(np.abs(want)))
    # dual basis: interpolating the element's own basis functions gives the identity
    phi = r.tabulate(0, e.points)[0]  # (npts, ndofs, vs) from the reference at the native points
    vals = phi.transpose(2, 0, 1).reshape(vs * e.points.shape[0], n)
    assert np.max(np.abs(e.interpolation_matrix @ vals - np.eye(n))) <= 1e-10


# ---- GPU ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("family,cell,degree,variant", [(N1E, TET, 3, LEG), (RT, TET, 4, LEG), (P, TET, 5, GLL), (RT, TRI, 2, LEG),
                                                        (P, HEX, 4, GLL), (N1E, TET, 1, LEG)])
def test_interpolate_batch_matches_numpy(ref, family, cell, degree, variant):
    import torch

    r, e = make_pair(ref, family, cell, degree, variant)
    M, npts, vs = r.interpolation_matrix, r.points.shape[0], r.value_size
    rng = np.random.default_rng(degree)
    for nc in (1, 15, 16, 17, 5003):
        v = rng.standard_normal((nc, vs * npts))
        want = v @ M.T
        scale = np.max(np.abs(want))
        got = e.interpolate_batch(v)  # host pointers, reference column order
        assert got.shape == (nc, r.dim) and np.max(np.abs(got - want)) <= 1e-12 * scale
        got = e.interpolate_batch(torch.from_numpy(v).cuda()).cpu().numpy()  # device pointers
        assert np.max(np.abs(got - want)) <= 1e-12 * scale
        # the order pull_back writes: values[c][p][j]
        vp = np.ascontiguousarray(v.reshape(nc, vs, npts).transpose(0, 2, 1)).reshape(nc, -1)
        got = e.interpolate_batch(vp, layout="point-major")
        assert np.max(np.abs(got - want)) <= 1e-12 * scale


@pytest.mark.gpu
def test_interpolate_batch_float32(ref):
    """float32 values in, float32 coefficients out (matrix and arithmetic stay float64)."""
    import torch

    r, e = make_pair(ref, 3, 3, 3, 11)
    M = r.interpolation_matrix
    v = np.random.default_rng(5).standard_normal((3001, M.shape[1])).astype(np.float32)
    want = v.astype(np.float64) @ M.T
    tol = 4e-7 * np.max(np.abs(want))
    got = e.interpolate_batch(v)
    assert got.dtype == np.float32 and np.max(np.abs(got - want)) <= tol
    got = e.interpolate_batch(torch.from_numpy(v).cuda(), layout="component-major")
    assert got.dtype == torch.float32 and np.max(np.abs(got.cpu().numpy() - want)) <= tol


@pytest.mark.gpu
def test_interpolate_after_pull_back_reproduces_coefficients(ref):
    """Interpolating a function of the element's own space returns its coefficients: tabulate at points(),
    combine with random coefficients, push forward to a physical cell, pull back, interpolate."""
    r, e = make_pair(ref, N1E, TET, 3, LEG)
    rng = np.random.default_rng(3)
    nc, npts, n = 257, r.points.shape[0], r.dim
    phi = e.tabulate(0, r.points)[0]  # (npts, ndofs, 3)
    coef = rng.standard_normal((nc, n))
    U = np.einsum("cn,pnj->cpj", coef, phi)  # reference values at the interpolation points
    J = np.eye(3)[None] + 0.2 * rng.standard_normal((nc, 3, 3))
    detJ, K = np.linalg.det(J), np.linalg.inv(J)
    u = e.push_forward(U, J, detJ, K)
    back = e.pull_back(u, J, detJ, K)  # (nc, npts, 3): point major
    got = e.interpolate_batch(back.reshape(nc, -1), layout="point-major")
    assert np.max(np.abs(got - coef)) <= 1e-10 * np.max(np.abs(coef))


@pytest.mark.gpu
def test_interpolate_errors(ref):
    import basix_b200 as bx

    r, e = make_pair(ref, N1E, TET, 1, LEG)
    with pytest.raises(RuntimeError):
        e.interpolate_batch(np.zeros((3, 5)))
    bare = bx.FiniteElement(cell_type=r.cell_type, polyset_type=r.polyset_type, degree=r.degree,
                            embedded_superdegree=r.embedded_superdegree, map_type=r.map_type, value_shape=r.value_shape,
                            coefficient_matrix=r.coefficient_matrix, entity_dofs=r.entity_dofs,
                            entity_transformations=r.entity_transformations)
    with pytest.raises(RuntimeError, match="without points"):
        bare.interpolate_batch(np.zeros((3, 18)))


# ---- compute_interpolation_operator (interpolation.cpp:16-116) on the device --------------------------------------------
OPERATOR_PAIRS = [
    ((1, 3, 2, 2), (1, 3, 3, 2)),      # P2 -> P3 tetrahedron (equal value sizes)
    ((1, 3, 3, 2), (1, 3, 1, 2)),      # P3 -> P1
    ((3, 3, 2, 11), (2, 3, 2, 11)),    # N1curl2 -> RT2 (vector to vector)
    ((2, 2, 2, 11), (3, 2, 3, 11)),    # RT2 -> N1curl3 triangle
    ((3, 3, 1, 11), (1, 3, 2, 2)),     # vector source, scalar target: one block per component
    ((1, 2, 3, 2), (2, 2, 2, 11)),     # scalar source, vector target
    ((1, 5, 2, 2), (1, 5, 1, 2)),      # Q2 -> Q1 hexahedron
    ((1, 4, 3, 2), (1, 4, 3, 1)),      # Q3 gll -> Q3 equispaced quadrilateral
]


@pytest.mark.gpu
@pytest.mark.parametrize("a0,a1", OPERATOR_PAIRS)
def test_compute_interpolation_operator(ref, a0, a1):
    """Same matrix as basix::compute_interpolation_operator on the same pair of (reference-built) elements."""
    import torch

    import basix_b200 as bx
    from conftest import make_pair

    r0, e0 = make_pair(ref, *a0)
    r1, e1 = make_pair(ref, *a1)
    want = ref.compute_interpolation_operator(r0, r1)
    got = bx.compute_interpolation_operator(e0, e1)
    assert got.shape == want.shape
    assert np.max(np.abs(got - want)) <= 1e-12 * max(1.0, np.max(np.abs(want)))
    dev = bx.compute_interpolation_operator(e0, e1, device=torch.device("cuda"))
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)


@pytest.mark.gpu
def test_compute_interpolation_operator_errors_and_native_elements(ref):
    import basix_b200 as bx
    from conftest import make_pair

    _, tri = make_pair(ref, 1, 2, 2, 2)
    _, tet = make_pair(ref, 1, 3, 2, 2)
    with pytest.raises(RuntimeError, match="different cell types"):
        bx.compute_interpolation_operator(tri, tet)
    # interpolating an element into itself gives the identity; natively built elements (Gauss-Jacobi points)
    for args in ((1, 3, 4, 2), (3, 3, 2, 11), (2, 2, 3, 11)):
        e = bx.create_element(*args)
        op = bx.compute_interpolation_operator(e, e)
        assert np.max(np.abs(op - np.eye(e.dim))) < 1e-11
    # P2 -> P3 of the native factory against the reference pair
    want = ref.compute_interpolation_operator(ref.RefElement(1, 3, 2, 2), ref.RefElement(1, 3, 3, 2))
    got = bx.compute_interpolation_operator(bx.create_element(1, 3, 2, 2), bx.create_element(1, 3, 3, 2))
    assert np.max(np.abs(got - want)) <= 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("family,cell,degree", [(3, 3, 3), (2, 3, 4), (3, 3, 1), (2, 3, 1), (3, 2, 2), (2, 2, 3), (3, 2, 1)])
@pytest.mark.parametrize("nc", [1, 16, 1000 + 13])
def test_interpolate_pulled_equals_pull_back_then_interpolate(family, cell, degree, nc):
    """pull_back fused into the interpolation (bx_interpolate_pulled_f64): the library's own pull_back followed by
    interpolate_batch to 1e-13 (another summation order in the contraction), host and device pointers bit-equal; and within
    1e-12 of the reference's pull_back times the matrix."""
    import torch

    import basix_b200 as bx
    from oracle import ref

    e = bx.create_element(family, cell, degree, 11)
    r = ref.RefElement(family, cell, degree, 11)
    tdim = 3 if cell == 3 else 2
    rng = np.random.default_rng(nc + degree)
    J = np.eye(tdim)[None] + 0.25 * rng.standard_normal((nc, tdim, tdim))
    detJ, K = np.linalg.det(J), np.ascontiguousarray(np.linalg.inv(J))
    npts = e.points.shape[0]
    u = rng.standard_normal((nc, npts, tdim))
    two = e.interpolate_batch(e.pull_back(u, J, detJ, K).reshape(nc, -1), "point-major")
    got = e.interpolate_pulled_batch(u, J, detJ, K)
    # fused multiply-adds in the map and another k order in the contraction: equal to rounding
    scale = max(np.max(np.abs(two)), 1.0)
    assert got.shape == (nc, e.dim) and np.max(np.abs(got - two)) <= 1e-13 * scale, np.max(np.abs(got - two)) / scale
    dev = e.interpolate_pulled_batch(*[torch.from_numpy(a).cuda() for a in (u, J, detJ, K)])
    assert np.array_equal(dev.cpu().numpy(), got)
    # the reference: its own pull_back, then our element's matrix in the reference's column order (the native element's
    # interpolation points differ from the reference element's, the map arithmetic does not)
    U = r.pull_back(u, J, detJ, K)  # (nc, npts, tdim)
    want = np.einsum("nk,ck->cn", e.interpolation_matrix, U.transpose(0, 2, 1).reshape(nc, -1))
    assert np.max(np.abs(got - want)) <= 1e-12 * max(np.max(np.abs(want)), 1.0)
    # float32 arrays: the float64 result to float32 accuracy
    got32 = e.interpolate_pulled_batch(*[a.astype(np.float32) for a in (u, J, detJ, K)])
    assert got32.dtype == np.float32 and np.allclose(got32, want, rtol=2e-4, atol=2e-4 * np.max(np.abs(want)))


@pytest.mark.gpu
def test_interpolate_pulled_rejects_what_it_does_not_fuse():
    import basix_b200 as bx

    p = bx.create_element(1, 3, 2, 2)
    with pytest.raises(RuntimeError):
        p.interpolate_pulled_batch(np.zeros((2, p.points.shape[0], 3)), np.zeros((2, 3, 3)), np.ones(2), np.zeros((2, 3, 3)))
