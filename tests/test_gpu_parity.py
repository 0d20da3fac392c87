"""Parity of the CUDA path (through the C ABI) against the reference oracle on identical inputs.

Bars (BASELINE.json north_star / SURVEY 8c):
  * permutations / integer indexing: bit-exact (array_equal);
  * polyset, tabulate, Piola maps, matrix DOF transformations: max|gpu-ref| <= 1e-12 * max|ref| per slab.
The polyset (simplex) and Piola kernels mirror the reference's operation order without FMA contraction and
are additionally checked for bit equality where that holds.
"""

import numpy as np
import pytest

from conftest import make_pair, random_points, slab_close

pytestmark = pytest.mark.gpu

P, RT, N1E = 1, 2, 3
INTERVAL, TRI, TET, QUAD, HEX = 1, 2, 3, 4, 5
GLL, EQ, LEG = 2, 1, 11


# ---- polyset::tabulate -------------------------------------------------------------------------------
@pytest.mark.parametrize("cell,degree,nd", [(INTERVAL, 0, 0), (INTERVAL, 1, 1), (INTERVAL, 7, 3), (TRI, 0, 2),
                                            (TRI, 1, 1), (TRI, 2, 1), (TRI, 6, 3), (TET, 0, 1), (TET, 1, 0),
                                            (TET, 3, 1), (TET, 5, 2), (TET, 4, 3), (QUAD, 0, 1), (QUAD, 3, 2),
                                            (HEX, 1, 0), (HEX, 4, 1), (HEX, 2, 2), (6, 0, 1), (6, 1, 0), (6, 3, 2),
                                            (6, 5, 1), (7, 0, 1), (7, 1, 2), (7, 4, 1), (7, 3, 3)])  # 6 prism, 7 pyramid
def test_polyset_tabulate(ref, cell, degree, nd):
    import basix_b200 as bx

    x = random_points(cell, 1003)
    if cell == 7:
        x[:2] = [[0.0, 0.0, 1.0], [0.25, 0.5, 0.5]]  # apex (special values) and a face point
    want = ref.tabulate_polynomial_set(cell, 0, degree, nd, x)
    got = bx.tabulate_polynomial_set(cell, 0, degree, nd, x)
    assert got.shape == want.shape
    slab_close(got, want, 1e-13)
    if cell in (INTERVAL, TRI, TET, 6, 7):
        # same operation order, no FMA contraction -> identical bits
        assert np.array_equal(got, want)


def test_polyset_golden_anchor():
    """SURVEY 8c anchors (reference output): triangle degree 2, nd=1 at (0.3, 0.2); hex degree 1."""
    import basix_b200 as bx

    got = bx.tabulate_polynomial_set(TRI, 0, 2, 1, np.array([[0.3, 0.2]]))[:, :, 0]
    want = np.array([[1.4142135623731, -0.8, -0.692820323027551, -0.489897948556636, 0, -1.42407864951343],
                     [0, 0, 6.92820323027551, 0, 0, -6.57267069006199],
                     [0, 6, 3.46410161513775, -9.79795897113271, -4.24264068711928, 1.09544511501033]])
    assert np.allclose(got, want, rtol=0, atol=5e-14)
    got = bx.tabulate_polynomial_set(HEX, 0, 1, 0, np.array([[.25, .5, .75]]))[0, :, 0]
    assert np.allclose(got, [1, 0.866025403784439, 0, 0, -0.866025403784439, -0.75, 0, 0], rtol=0, atol=5e-15)


def test_polyset_device_pointers(ref):
    import torch

    import basix_b200 as bx

    x = random_points(TET, 4099)
    want = ref.tabulate_polynomial_set(TET, 0, 4, 2, x)
    got = bx.tabulate_polynomial_set(TET, 0, 4, 2, torch.from_numpy(x).cuda())
    assert got.is_cuda
    assert np.array_equal(got.cpu().numpy(), want)


# ---- FiniteElement::tabulate ---------------------------------------------------------------------------
ELEMENTS = [
    (P, TRI, 1, GLL), (P, TRI, 4, GLL), (P, TET, 1, GLL), (P, TET, 2, EQ), (P, TET, 5, GLL), (P, INTERVAL, 3, GLL),
    (P, QUAD, 2, GLL), (P, HEX, 1, GLL), (P, HEX, 4, GLL), (N1E, TET, 3, LEG), (RT, TET, 4, LEG), (RT, TRI, 2, LEG),
    (N1E, TRI, 1, LEG),
]


@pytest.mark.parametrize("family,cell,degree,variant", ELEMENTS)
@pytest.mark.parametrize("nd", [0, 1, 2])
def test_tabulate(ref, family, cell, degree, variant, nd):
    r, e = make_pair(ref, family, cell, degree, variant)
    x = random_points(cell, 777, seed=degree + 10 * nd)
    want = r.tabulate(nd, x)
    got = e.tabulate(nd, x)
    assert got.shape == want.shape == e.tabulate_shape(nd, len(x))
    slab_close(got, want, 1e-12)


def test_tabulate_golden_p1_triangle(ref):
    r, e = make_pair(ref, P, TRI, 1, GLL)
    got = e.tabulate(1, np.array([[0.3, 0.2]])).reshape(3, 3)
    assert np.allclose(got, [[.5, .3, .2], [-1, 1, 0], [-1, 0, 1]], rtol=0, atol=1e-15)


def test_tabulate_dof_ordering(ref):
    """dof_ordering scatter, finite-element.cpp:1880-1886 (reference test: test_dof_ordering.py:21-57)."""
    order = [3, 0, 5, 1, 4, 2]
    r0, e0 = make_pair(ref, P, TRI, 2, GLL)
    r1, e1 = make_pair(ref, P, TRI, 2, GLL, dof_ordering=order)
    x = random_points(TRI, 200)
    t0, t1 = e0.tabulate(1, x), e1.tabulate(1, x)
    slab_close(t1, r1.tabulate(1, x), 1e-12)
    for i, j in enumerate(order):
        assert np.array_equal(t0[:, :, i, :], t1[:, :, j, :])


def test_tabulate_ragged_and_device(ref):
    import torch

    r, e = make_pair(ref, P, TET, 3, GLL)
    for npts in (1, 7, 31, 33, 257):
        x = random_points(TET, npts, seed=npts)
        slab_close(e.tabulate(1, x), r.tabulate(1, x), 1e-12)
        xd = torch.from_numpy(x).cuda()
        got = e.tabulate(1, xd)
        assert got.is_cuda
        slab_close(got.cpu().numpy(), r.tabulate(1, x), 1e-12)
    assert e.tabulate(1, np.zeros((0, 3))).shape == (4, 0, 20, 1)


def test_tabulate_wrong_point_dim(ref):
    r, e = make_pair(ref, P, TET, 2, EQ)
    with pytest.raises(RuntimeError, match=r"Point dim \(2\) does not match element dim \(3\)\."):
        e.tabulate(0, np.zeros((4, 2)))


def test_tabulate_partition_of_unity_large(ref):
    """Size-independent property at a size the oracle is too slow for: Lagrange basis sums to 1 and
    derivative slabs sum to 0 (reference test_lagrange.py:857-864)."""
    import torch

    r, e = make_pair(ref, P, TET, 5, GLL)
    x = torch.from_numpy(random_points(TET, 300_000, seed=5)).cuda()
    t = e.tabulate(2, x)
    s = t.sum(dim=2)[..., 0]
    assert float((s[0] - 1).abs().max()) < 1e-11
    scale = float(t.abs().max())
    assert float(s[1:].abs().max()) < 1e-11 * scale


# ---- push_forward / pull_back ----------------------------------------------------------------------------
def jacobians(nc, gdim, tdim, seed=7):
    rng = np.random.default_rng(seed)
    J = np.zeros((nc, gdim, tdim))
    J[:, :tdim, :tdim] = np.eye(tdim)
    J += 0.3 * rng.standard_normal((nc, gdim, tdim))
    if gdim == tdim:
        detJ = np.linalg.det(J)
        K = np.linalg.inv(J)
    else:
        JtJ = np.einsum("cji,cjk->cik", J, J)
        detJ = np.sqrt(np.linalg.det(JtJ))
        K = np.einsum("cij,ckj->cik", np.linalg.inv(JtJ), J)
    return np.ascontiguousarray(J), np.ascontiguousarray(detJ), np.ascontiguousarray(K)


MAP_ELEMENTS = [(P, TET, 2, EQ, 1), (N1E, TET, 3, LEG, 3), (RT, TET, 4, LEG, 3), (N1E, TRI, 2, LEG, 2),
                (RT, TRI, 2, LEG, 2), (7, TET, 1, 0, 9), (11, TRI, 1, 0, 4), (7, TRI, 2, 0, 4), (11, TET, 1, 0, 9)]


@pytest.mark.parametrize("family,cell,degree,variant,vs", MAP_ELEMENTS)
@pytest.mark.parametrize("rows", [1, 5])
def test_push_pull_square(ref, family, cell, degree, variant, vs, rows):
    r, e = make_pair(ref, family, cell, degree, variant)
    tdim = 3 if cell == TET else 2
    nc = 1234
    J, detJ, K = jacobians(nc, tdim, tdim)
    U = np.random.default_rng(1).standard_normal((nc, rows, vs))
    want = r.push_forward(U, J, detJ, K)
    got = e.push_forward(U, J, detJ, K)
    assert got.shape == want.shape
    assert np.array_equal(got, want), np.max(np.abs(got - want))
    back_ref = r.pull_back(want, J, detJ, K)
    back = e.pull_back(got, J, detJ, K)
    assert np.array_equal(back, back_ref)
    assert np.allclose(back, U, rtol=1e-9, atol=1e-9)  # reference test_mappings.py round trip


@pytest.mark.parametrize("family,variant,vs", [(N1E, LEG, 2), (RT, LEG, 2), (7, 0, 4), (11, 0, 4), (P, GLL, 1)])
def test_push_pull_manifold(ref, family, variant, vs):
    """2-D cells embedded in 3-D: non-square J (reference test_mappings.py:66-81)."""
    r, e = make_pair(ref, family, TRI, 2, variant)
    nc = 301
    J, detJ, K = jacobians(nc, 3, 2)
    U = np.random.default_rng(2).standard_normal((nc, 4, vs))
    want = r.push_forward(U, J, detJ, K)
    got = e.push_forward(U, J, detJ, K)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(e.pull_back(got, J, detJ, K), r.pull_back(want, J, detJ, K))


@pytest.mark.parametrize("family,cell,gdim,vs", [(N1E, TET, 3, 3), (RT, TET, 3, 3), (7, TET, 3, 9), (11, TET, 3, 9),
                                                (N1E, TRI, 3, 2), (RT, TRI, 3, 2), (RT, TRI, 2, 2), (11, TRI, 3, 4)])
@pytest.mark.parametrize("rows,nc", [(2, 1003), (3, 77), (31, 130), (32, 65), (33, 129), (64, 33), (100, 41), (700, 5)])
def test_push_pull_rows_around_the_warp_size(ref, family, cell, gdim, vs, rows, nc):
    """Several rows per cell: the warp-staged kernel (a warp owns 32 rows of the flat (cell, row) space and stages the
    matrices of the cells they belong to) for row counts below, at and above the warp size, totals that are not a
    multiple of 32, square and non-square J, all four Piola maps -- bit for bit against the reference."""
    degree = 1 if family in (7, 11) else 2
    r, e = make_pair(ref, family, cell, degree, 0 if family in (7, 11) else LEG)
    tdim = 3 if cell == TET else 2
    J, detJ, K = jacobians(nc, gdim, tdim, seed=rows)
    U = np.random.default_rng(rows + nc).standard_normal((nc, rows, vs))
    want = r.push_forward(U, J, detJ, K)
    got = e.push_forward(U, J, detJ, K)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(e.pull_back(got, J, detJ, K), r.pull_back(want, J, detJ, K))


@pytest.mark.parametrize("family,cell,degree,variant,vs", [(N1E, TET, 3, LEG, 3), (RT, TET, 4, LEG, 3), (RT, TRI, 2, LEG, 2),
                                                          (P, TET, 2, EQ, 1), (7, TET, 1, 0, 9)])
def test_push_forward_broadcast(ref, family, cell, degree, variant, vs):
    """One slab of reference values for all cells == the reference's push_forward on the replicated array
    (bit for bit: same kernel arithmetic); host and device pointers, float64 and float32."""
    import torch

    r, e = make_pair(ref, family, cell, degree, variant)
    tdim = 3 if cell == TET else 2
    rng = np.random.default_rng(degree)
    for nc, rows in ((1, 1), (37, 5), (1201, 45)):
        J, detJ, K = jacobians(nc, tdim, tdim, seed=nc)
        U0 = rng.standard_normal((rows, vs))
        want = r.push_forward(np.ascontiguousarray(np.broadcast_to(U0, (nc, rows, vs))), J, detJ, K)
        got = e.push_forward_broadcast(U0, J, detJ, K)
        assert np.array_equal(got, want)
        dev = [torch.from_numpy(a).cuda() for a in (U0, J, detJ, K)]
        assert np.array_equal(e.push_forward_broadcast(*dev).cpu().numpy(), want)
        got32 = e.push_forward_broadcast(*[a.astype(np.float32) for a in (U0, J, detJ, K)])
        assert got32.dtype == np.float32 and np.allclose(got32, want, rtol=2e-4, atol=2e-4 * np.max(np.abs(want)))


def test_map_device_pointers_and_golden(ref):
    import torch

    r, e = make_pair(ref, N1E, TET, 3, LEG)
    J = np.array([[[2, .5, .25], [.125, 1.5, .75], [.0625, .3125, 1]]])
    detJ, K = np.linalg.det(J), np.linalg.inv(J)
    U = np.array([[[1., 2, 3]]])
    got = e.push_forward(U, J, detJ, K)
    assert np.allclose(got, [[[0.38140267927502, 0.7123719464145, 2.37037037037037]]], rtol=0, atol=1e-14)
    assert np.allclose(e.pull_back(U, J, detJ, K), [[[2.4375, 4.4375, 4.75]]], rtol=0, atol=1e-14)
    nc = 5000
    Jb, db, Kb = jacobians(nc, 3, 3)
    Ub = np.random.default_rng(3).standard_normal((nc, 45, 3))
    t = [torch.from_numpy(a).cuda() for a in (Ub, Jb, db, Kb)]
    got = e.push_forward(*t)
    assert got.is_cuda and np.array_equal(got.cpu().numpy(), r.push_forward(Ub, Jb, db, Kb))


# ---- DOF transformations -----------------------------------------------------------------------------------
def cell_infos(nc, seed=11):
    rng = np.random.default_rng(seed)
    ci = rng.integers(0, 2 ** 30, nc).astype(np.uint32)
    ci[:4] = [0, 1, 0x2ABCD, 0x3FFFFFFF]
    return ci


PERM_ELEMENTS = [(P, TRI, 3, GLL), (P, TET, 5, GLL), (P, TET, 3, EQ), (P, HEX, 3, GLL), (P, QUAD, 4, GLL),
                 (P, HEX, 4, GLL)]
ALL_OPS = ["T_apply", "Tt_apply", "Tt_inv_apply", "Tinv_apply", "Tt_apply_right", "Tinv_apply_right",
           "T_apply_right", "Tt_inv_apply_right"]


@pytest.mark.parametrize("family,cell,degree,variant", PERM_ELEMENTS)
def test_permute_bit_exact(ref, family, cell, degree, variant):
    r, e = make_pair(ref, family, cell, degree, variant)
    nc = 2049
    ci = cell_infos(nc)
    d = np.tile(np.arange(r.dim, dtype=np.int32), (nc, 1)) + 1000 * np.arange(nc, dtype=np.int32)[:, None]
    for inverse in (False, True):
        want = r.permute_batch(d.copy(), ci, inverse=inverse)
        got = d.copy()
        e.permute_batch(got, ci, inverse=inverse)
        assert np.array_equal(got, want)
    # round trip
    rt = d.copy()
    e.permute_batch(rt, ci)
    e.permute_batch(rt, ci, inverse=True)
    assert np.array_equal(rt, d)


def test_permute_golden_p5_tet(ref):
    """SURVEY 8c anchors for P5 tetrahedron."""
    r, e = make_pair(ref, P, TET, 5, GLL)
    d = np.arange(56, dtype=np.int32)
    e.permute(d, 0x2ABCD)
    want = [0, 1, 2, 3, 4, 5, 6, 7, 11, 10, 9, 8, 12, 13, 14, 15, 19, 18, 17, 16, 20, 21, 22, 23, 27, 26, 25, 24, 33,
            32, 30, 31, 29, 28, 34, 37, 39, 35, 38, 36, 40, 43, 45, 41, 44, 42, 51, 50, 48, 49, 47, 46, 52, 53, 54, 55]
    assert d.tolist() == want
    d = np.arange(56, dtype=np.int32)
    e.permute(d, 0x1000)
    assert d[4:8].tolist() == [7, 6, 5, 4]
    d = np.arange(56, dtype=np.int32)
    e.permute(d, 0x1)
    assert d[28:34].tolist() == [28, 31, 33, 29, 32, 30]


@pytest.mark.parametrize("family,cell,degree,variant", PERM_ELEMENTS[:4])
@pytest.mark.parametrize("dtype", [np.float64, np.float32, np.int32])
@pytest.mark.parametrize("bs", [1, 3])
def test_T_apply_permutation_elements_bit_exact(ref, family, cell, degree, variant, dtype, bs):
    r, e = make_pair(ref, family, cell, degree, variant)
    nc = 515
    ci = cell_infos(nc)
    rng = np.random.default_rng(5)
    data = (rng.standard_normal((nc, r.dim * bs)) * 100).astype(dtype)
    ops = ALL_OPS if dtype != np.int32 else ALL_OPS[:4]
    for op in ops:
        want = r.T_apply_batch(op, data.copy(), bs, ci)
        got = data.copy()
        e.T_apply_batch(got, bs, ci, op=op)
        assert np.array_equal(got, want), op


MATRIX_ELEMENTS = [(RT, TET, 4, LEG), (N1E, TET, 3, LEG), (RT, TET, 2, LEG), (N1E, TRI, 3, LEG), (N1E, HEX, 2, LEG)]


@pytest.mark.parametrize("family,cell,degree,variant", MATRIX_ELEMENTS)
@pytest.mark.parametrize("bs", [1, 3, 16])
def test_T_apply_matrix_elements(ref, family, cell, degree, variant, bs):
    r, e = make_pair(ref, family, cell, degree, variant)
    nc = 300
    ci = cell_infos(nc)
    data = np.random.default_rng(6).standard_normal((nc, r.dim * bs))
    for op in ALL_OPS:
        want = r.T_apply_batch(op, data.copy(), bs, ci)
        got = data.copy()
        e.T_apply_batch(got, bs, ci, op=op)
        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), op


def test_T_apply_direct_kernel():
    """The register form of the matrix kernel (dof_transform_direct.cuh) is a development alternative selected with
    BX_DOF_TUNE="direct=1"; the knob is read once per process, so it is checked in a child process against the
    default tile kernel of this process's reference run."""
    import os
    import subprocess
    import sys

    code = (
        "import numpy as np, sys\n"
        "sys.path.insert(0, %r)\n"
        "import basix_b200 as bx\n"
        "from oracle import ref\n"
        "n0 = bx.kernel_launch_count()\n"
        "for fam, deg in ((2, 4), (3, 3), (2, 2)):\n"
        "    r, e = ref.RefElement(fam, 3, deg, 11), bx.create_element(fam, 3, deg, 11)\n"
        "    rng = np.random.default_rng(7)\n"
        "    ci = rng.integers(0, 2**30, 4001).astype(np.uint32)\n"
        "    data = rng.standard_normal((4001, r.dim))\n"
        "    for op in ('T_apply', 'Tt_apply', 'Tinv_apply', 'Tt_inv_apply'):\n"
        "        want = data.copy()\n"
        "        r.T_apply_batch(op, want, 1, ci)\n"
        "        got = data.copy()\n"
        "        e.T_apply_batch(got, 1, ci, op=op)\n"
        "        assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want)), (fam, deg, op)\n"
        "assert bx.kernel_launch_count() > n0\n"
        "print('ok')\n"
    ) % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, BX_DOF_TUNE="direct=1")
    out = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_T_apply_single_cell_reference_signature(ref):
    """The reference API: one cell per call (finite_element.py:153-166)."""
    r, e = make_pair(ref, RT, TET, 2, LEG)
    for ci in (0x1, 0x2, 0x4, 0x5A7):
        want = r.T_apply(np.arange(1, 16, dtype=np.float64), 1, ci)
        got = np.arange(1, 16, dtype=np.float64)
        e.T_apply(got, 1, ci)
        assert np.allclose(got, want, rtol=0, atol=1e-13)


def test_T_apply_inverse_pairs(ref):
    """T^-1 T = I and (T^t)^-1 T^t = I on random data (reference test_dof_transformations.py:45-66 style)."""
    import torch

    r, e = make_pair(ref, RT, TET, 4, LEG)
    nc = 20000
    ci = torch.from_numpy(cell_infos(nc).astype(np.int64)).cuda().to(torch.uint32)
    data = torch.randn(nc, 70 * 2, dtype=torch.float64, device="cuda")
    for a, b in (("T_apply", "Tinv_apply"), ("Tt_apply", "Tt_inv_apply"), ("Tt_apply_right", "Tt_inv_apply_right")):
        w = data.clone()
        e.T_apply_batch(w, 2, ci, op=a)
        assert not torch.equal(w, data)
        e.T_apply_batch(w, 2, ci, op=b)
        assert float((w - data).abs().max()) < 1e-11


def test_errors(ref):
    r, e = make_pair(ref, RT, TET, 2, LEG)
    with pytest.raises(RuntimeError, match="not permutations"):
        e.permute(np.arange(15, dtype=np.int32), 1)
    with pytest.raises(RuntimeError):
        e.T_apply_batch(np.zeros((2, 14)), 1, np.zeros(2, dtype=np.uint32))


# ---- fused simplex kernel (tabulate_fused.cuh) ------------------------------------------------------------
FUSED = [(TET, 2, 3), (TET, 3, 3), (TET, 4, 2), (TET, 5, 2), (TET, 5, 1), (TET, 5, 0), (TET, 6, 1), (TET, 6, 0),
         (TRI, 3, 3), (TRI, 5, 2), (TRI, 8, 2), (TRI, 7, 1), (TRI, 6, 0), (INTERVAL, 7, 2), (INTERVAL, 8, 0)]


@pytest.mark.parametrize("cell,degree,nd", FUSED)
def test_tabulate_fused_path(ref, cell, degree, nd):
    import basix_b200 as bx

    r, e = make_pair(ref, P, cell, degree, GLL)
    with bx.tabulate_paths("fused"):
        assert e.tabulate_path(nd) == "fused"
        # more tiles than CTAs (ring wraps many times), ragged tail
        for npts in (5, 32, 20011):
            x = random_points(cell, npts, seed=npts + degree)
            slab_close(e.tabulate(nd, x), r.tabulate(nd, x, chunk=4096, nthreads=8), 1e-12)


DOP = [c for c in FUSED if c != (TET, 6, 1)] + [(TET, 1, 2), (TRI, 2, 3), (TRI, 8, 3), (TRI, 6, 4), (INTERVAL, 10, 4), (INTERVAL, 1, 2)]


@pytest.mark.parametrize("cell,degree,nd", DOP)
def test_tabulate_derivative_operator_path(ref, cell, degree, nd):
    """Every slab from the values of the orthonormal set (tabulate_dop.cu); also nd above the degree (zero slabs)."""
    import basix_b200 as bx

    r, e = make_pair(ref, P, cell, degree, GLL)
    with bx.tabulate_paths("derivative-operator"):
        assert e.tabulate_path(nd) == "derivative-operator"
        for npts in (1, 63, 64, 65, 20011):
            x = random_points(cell, npts, seed=npts + degree)
            slab_close(e.tabulate(nd, x), r.tabulate(nd, x, chunk=4096, nthreads=8), 1e-12)


def test_tabulate_operands_too_large_for_shared_memory_fall_back(ref):
    """P6 on the tetrahedron with derivatives: the resident operands exceed shared memory -> ring-streamed fused kernel."""
    r, e = make_pair(ref, P, TET, 6, GLL)
    assert e.tabulate_path(0) == "derivative-operator" and e.tabulate_path(1) == "fused"
    x = random_points(TET, 4001, seed=1)
    slab_close(e.tabulate(1, x), r.tabulate(1, x, chunk=4096, nthreads=8), 1e-12)


@pytest.mark.parametrize("family,cell,degree,variant,nd", [(N1E, TET, 3, LEG, 1), (RT, TET, 4, LEG, 1), (N1E, TRI, 2, LEG, 2),
                                                           (RT, TRI, 3, LEG, 0)])
def test_tabulate_derivative_operator_vector_valued(ref, family, cell, degree, variant, nd):
    import basix_b200 as bx

    r, e = make_pair(ref, family, cell, degree, variant)
    with bx.tabulate_paths("derivative-operator"):
        assert e.tabulate_path(nd) == "derivative-operator"
        x = random_points(cell, 5003, seed=degree)
        slab_close(e.tabulate(nd, x), r.tabulate(nd, x), 1e-12)


def test_tabulate_fused_dof_ordering_and_device(ref):
    import torch

    order = list(np.random.default_rng(0).permutation(35))
    import basix_b200 as bx

    r, e = make_pair(ref, P, TET, 4, GLL, dof_ordering=[int(v) for v in order])
    x = random_points(TET, 3333, seed=9)
    want = r.tabulate(1, x)
    for path in ("fused", "derivative-operator"):
        with bx.tabulate_paths(path):
            assert e.tabulate_path(1) == path
            slab_close(e.tabulate(1, x), want, 1e-12)
            got = e.tabulate(1, torch.from_numpy(x).cuda())
            slab_close(got.cpu().numpy(), want, 1e-12)


# ---- tensor-product kernel (tabulate_tp.cu) -----------------------------------------------------------------
TP = [(P, QUAD, 1, GLL), (P, QUAD, 3, GLL), (P, QUAD, 5, EQ), (P, HEX, 1, GLL), (P, HEX, 2, GLL), (P, HEX, 3, EQ),
      (P, HEX, 4, GLL), (P, HEX, 6, GLL)]


@pytest.mark.parametrize("family,cell,degree,variant", TP)
@pytest.mark.parametrize("nd", [0, 1, 2])
def test_tabulate_tensor_product_path(ref, family, cell, degree, variant, nd):
    """Q elements on quad/hex (reference test_tensor_products.py:60-200 checks the same structure)."""
    r, e = make_pair(ref, family, cell, degree, variant)
    assert e.tabulate_path(nd) == "tensor-product"
    for npts in (1, 33, 4099):
        x = random_points(cell, npts, seed=npts + degree)
        slab_close(e.tabulate(nd, x), r.tabulate(nd, x), 1e-12)


@pytest.mark.parametrize("cell,degree", [(QUAD, 1), (QUAD, 4), (QUAD, 6), (HEX, 1), (HEX, 2), (HEX, 3), (HEX, 4)])
@pytest.mark.parametrize("nd", [0, 1, 2])
def test_tabulate_tensor_product_staged_and_unstaged_kernels(ref, cell, degree, nd):
    """Both tensor-product kernels (tabulate_tp_staged.cu: one slab of a 32-point tile at a time through shared
    memory; tabulate_tp.cu: direct stores) against the reference: ragged tails, odd and even point counts (16- and
    8-byte copy-out), odd and even column counts (dense and padded tile pitch)."""
    import basix_b200 as bx

    r, e = make_pair(ref, P, cell, degree, GLL if degree > 2 else 0)
    for allowed in (("tensor-product",), ("tensor-product", "tensor-product-staged")):
        with bx.tabulate_paths(*allowed):
            assert e.tabulate_path(nd) == "tensor-product"
            for npts in (1, 31, 32, 33, 64, 4098, 20011):
                x = random_points(cell, npts, seed=npts + degree)
                slab_close(e.tabulate(nd, x), r.tabulate(nd, x), 1e-12)


def test_tabulate_non_tensor_product_elements_take_generic_path(ref):
    """Serendipity on a quadrilateral is not a tensor product: the rank-1 test must reject it."""
    r, e = make_pair(ref, 10, QUAD, 3, LEG, 7)
    assert e.tabulate_path(1) == "generic"
    x = random_points(QUAD, 500)
    slab_close(e.tabulate(1, x), r.tabulate(1, x), 1e-12)


def test_tabulate_tp_dof_ordering(ref):
    order = [int(v) for v in np.random.default_rng(1).permutation(27)]
    r, e = make_pair(ref, P, HEX, 2, GLL, dof_ordering=order)
    assert e.tabulate_path(1) == "tensor-product"
    x = random_points(HEX, 1000)
    slab_close(e.tabulate(1, x), r.tabulate(1, x), 1e-12)


# ---- register path for low-degree simplex elements (tabulate_small.cu) --------------------------------------
SMALL = [(P, TRI, 1, GLL, 1), (P, TRI, 1, GLL, 3), (P, TRI, 2, GLL, 2), (P, TET, 1, GLL, 2), (P, TET, 2, EQ, 1),
         (P, INTERVAL, 2, GLL, 3), (N1E, TRI, 1, LEG, 1), (N1E, TET, 1, LEG, 1), (RT, TRI, 2, LEG, 1), (RT, TET, 1, LEG, 0)]


@pytest.mark.parametrize("family,cell,degree,variant,nd", SMALL)
def test_tabulate_register_path(ref, family, cell, degree, variant, nd):
    r, e = make_pair(ref, family, cell, degree, variant)
    assert e.tabulate_path(nd) == "register"
    for npts in (1, 127, 128, 129, 50021):  # one ragged tile, exact tile, tile + 1, many tiles per CTA
        x = random_points(cell, npts, seed=npts + degree)
        got = e.tabulate(nd, x)
        slab_close(got, r.tabulate(nd, x), 1e-12)


def test_tabulate_register_path_dof_ordering_and_device(ref):
    import torch

    order = [int(v) for v in np.random.default_rng(2).permutation(6)]
    r, e = make_pair(ref, P, TRI, 2, GLL, dof_ordering=order)
    assert e.tabulate_path(1) == "register"
    x = random_points(TRI, 7777, seed=3)
    want = r.tabulate(1, x)
    slab_close(e.tabulate(1, x), want, 1e-12)
    slab_close(e.tabulate(1, torch.from_numpy(x).cuda()).cpu().numpy(), want, 1e-12)


# ---- DOF transformations: both device implementations -----------------------------------------------------
@pytest.mark.parametrize("family,cell,degree,variant", [(P, TET, 5, GLL), (RT, TET, 4, LEG), (N1E, TET, 3, LEG),
                                                         (P, HEX, 4, GLL), (N1E, HEX, 2, LEG)])
@pytest.mark.parametrize("offset", [0, 1])
def test_T_apply_aligned_and_unaligned_buffers(ref, family, cell, degree, variant, offset):
    """offset 0: 16-byte aligned rows -> the TMA kernels (when the cell pitch allows); offset 1: the same data one
    double further into the allocation -> the register-staged fallback.  Both must match the reference."""
    import torch

    r, e = make_pair(ref, family, cell, degree, variant)
    nc = 3001  # several tiles, ragged last tile
    ci = cell_infos(nc, seed=degree)
    for bs in (1, 2):
        data = np.random.default_rng(bs).standard_normal((nc, r.dim * bs))
        want = r.T_apply_batch("T_apply", data.copy(), bs, ci)
        flat = torch.empty(data.size + 2, dtype=torch.float64, device="cuda")
        view = flat[offset:offset + data.size].view(nc, r.dim * bs)
        view.copy_(torch.from_numpy(data))
        e.T_apply_batch(view, bs, torch.from_numpy(ci.view(np.int32)).cuda())
        got = view.cpu().numpy()
        if r.dof_transformations_are_permutations:
            assert np.array_equal(got, want)
        else:
            assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
        assert flat[0].item() == flat[0].item()  # no stray writes trip the allocator's guard in later tests


@pytest.mark.parametrize("op", ["T_apply", "Tt_apply", "Tt_inv_apply", "Tinv_apply", "Tt_apply_right",
                                "Tinv_apply_right", "T_apply_right", "Tt_inv_apply_right"])
def test_all_eight_operators_large_batch(ref, op):
    """Every T flavour (finite-element.h:1839-1998) over a batch large enough for many tiles per CTA."""
    nc = 20011
    ci = cell_infos(nc, seed=17)
    for family, cell, degree, variant in [(RT, TET, 4, LEG), (P, TET, 5, GLL)]:
        r, e = make_pair(ref, family, cell, degree, variant)
        for bs in (1, 3):
            data = np.random.default_rng(bs).standard_normal((nc, r.dim * bs))
            want = r.T_apply_batch(op, data.copy(), bs, ci, nthreads=8)
            got = data.copy()
            e.T_apply_batch(got, bs, ci, op=op)
            if r.dof_transformations_are_permutations:
                assert np.array_equal(got, want)
            else:
                assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


# ---- natively constructed elements (csrc/create.cu) through the whole device path ---------------------------
@pytest.mark.parametrize("cell,degree,variant,nd", [(TRI, 1, GLL, 1), (TET, 5, GLL, 2), (HEX, 4, GLL, 1), (QUAD, 3, EQ, 2),
                                                    (TET, 3, EQ, 1), (INTERVAL, 6, GLL, 2)])
def test_native_create_element_tabulate_and_permute(ref, cell, degree, variant, nd):
    import basix_b200 as bx

    r = ref.RefElement(P, cell, degree, variant)
    e = bx.create_element(bx.ElementFamily.P, cell, degree, variant)
    x = random_points(cell, 2049, seed=degree)
    slab_close(e.tabulate(nd, x), r.tabulate(nd, x), 1e-12)
    if not r.dof_transformations_are_identity:
        ci = cell_infos(257, seed=degree)
        d = np.tile(np.arange(r.dim, dtype=np.int32), (len(ci), 1))
        want = r.permute_batch(d.copy(), ci)
        e.permute_batch(d, ci)
        assert np.array_equal(d, want)


# ---- float32 API (the reference's FiniteElement<float>, finite-element.cpp:2051) ----------------------------
F32_TAB = [(P, TRI, 1, GLL, 1), (P, TET, 2, EQ, 1), (P, TET, 5, GLL, 2), (P, HEX, 4, GLL, 1), (P, QUAD, 3, GLL, 2),
           (N1E, TET, 3, LEG, 1), (RT, TET, 2, LEG, 0)]


@pytest.mark.parametrize("family,cell,degree,variant,nd", F32_TAB)
def test_tabulate_float32(ref, family, cell, degree, variant, nd):
    """Every tabulate path with float32 points / table: compared with the float64 reference evaluated at the
    same (float-representable) points, to a few float32 ulp of the slab maximum."""
    import torch

    r, e = make_pair(ref, family, cell, degree, variant)
    for npts in (1, 129, 20011):
        x32 = random_points(cell, npts, seed=npts).astype(np.float32)
        want = r.tabulate(nd, x32.astype(np.float64))
        got = e.tabulate(nd, x32)
        assert got.dtype == np.float32 and got.shape == want.shape
        slab_close(got.astype(np.float64), want, 1e-6)
        got_dev = e.tabulate(nd, torch.from_numpy(x32).cuda())
        assert got_dev.dtype == torch.float32 and np.array_equal(got_dev.cpu().numpy(), got)


@pytest.mark.parametrize("family,cell,degree,variant,vs", MAP_ELEMENTS)
def test_push_pull_float32(ref, family, cell, degree, variant, vs):
    r, e = make_pair(ref, family, cell, degree, variant)
    tdim = {TRI: 2, TET: 3, QUAD: 2, HEX: 3, INTERVAL: 1}[cell]
    nc, rows = 513, 3
    J, detJ, K = jacobians(nc, tdim, tdim)
    U = np.random.default_rng(3).standard_normal((nc, rows, vs))
    a32 = [a.astype(np.float32) for a in (U, J, detJ, K)]
    a64 = [a.astype(np.float64) for a in a32]
    u = e.push_forward(*a32)
    want = r.push_forward(*a64)
    assert u.dtype == np.float32 and u.shape == want.shape
    assert np.max(np.abs(u - want)) <= 2e-5 * np.max(np.abs(want))
    back = e.pull_back(u, *a32[1:])
    want_back = r.pull_back(u.astype(np.float64), *a64[1:])
    # float arithmetic through K^T (.) K / J (.) J^T: the double Piola maps square the conditioning of J
    tol = 1e-3 if vs >= 4 else 1e-4
    assert back.dtype == np.float32 and np.max(np.abs(back - want_back)) <= tol * np.max(np.abs(want_back))


# ---- float32 against the reference's OWN float instantiation (FiniteElement<float>, oracle/ref_shim.cpp bxref32_*) ----
@pytest.mark.parametrize("family,cell,degree,variant,nd", F32_TAB)
def test_tabulate_float32_vs_reference_float(ref, family, cell, degree, variant, nd):
    """The library narrows a float64 evaluation, the reference's FiniteElement<float> computes in float: the two agree
    to within the rounding error of the reference's float path (measured here against the reference's double path at
    the same points), and the library's table is the more accurate of the two."""
    r64, e = make_pair(ref, family, cell, degree, variant)
    r32 = ref.RefElement32(family, cell, degree, variant)
    x32 = random_points(cell, 4099, seed=7).astype(np.float32)
    exact = r64.tabulate(nd, x32.astype(np.float64))
    want = r32.tabulate(nd, x32)
    got = e.tabulate(nd, x32)
    assert got.dtype == want.dtype == np.float32 and got.shape == want.shape
    eps = float(np.finfo(np.float32).eps)
    for d in range(want.shape[0]):
        scale = np.max(np.abs(exact[d]))
        ref_err = np.max(np.abs(want[d].astype(np.float64) - exact[d]))
        our_err = np.max(np.abs(got[d].astype(np.float64) - exact[d]))
        assert our_err <= eps * scale, (d, our_err, scale)
        assert np.max(np.abs(got[d].astype(np.float64) - want[d])) <= ref_err + eps * scale, (d, ref_err, scale)


@pytest.mark.parametrize("family,cell,degree,variant,vs", MAP_ELEMENTS)
def test_push_pull_float32_vs_reference_float_bit_exact(ref, family, cell, degree, variant, vs):
    """The float32 maps run the reference's operation order in float arithmetic: identical bits to FiniteElement<float>."""
    _, e = make_pair(ref, family, cell, degree, variant)
    r32 = ref.RefElement32(family, cell, degree, variant)
    tdim = {TRI: 2, TET: 3}[cell]
    nc, rows = 2049, 3
    J, detJ, K = jacobians(nc, tdim, tdim)
    U = np.random.default_rng(3).standard_normal((nc, rows, vs))
    a32 = [a.astype(np.float32) for a in (U, J, detJ, K)]
    u = e.push_forward(*a32)
    want = r32.push_forward(*a32)
    assert u.dtype == np.float32 and np.array_equal(u, want)
    assert np.array_equal(e.pull_back(u, *a32[1:]), r32.pull_back(want, *a32[1:]))


@pytest.mark.parametrize("family,cell,degree,variant", [(RT, TET, 4, LEG), (N1E, TET, 3, LEG), (P, TET, 5, GLL)])
@pytest.mark.parametrize("op", ["T_apply", "Tinv_apply", "Tt_apply_right"])
def test_T_apply_float32_vs_reference_float(ref, family, cell, degree, variant, op):
    """float32 data: against FiniteElement<float> (float LU sweeps) to float rounding; permutation elements bit for bit."""
    r64, e = make_pair(ref, family, cell, degree, variant)
    r32 = ref.RefElement32(family, cell, degree, variant)
    nc, bs = 3001, 2
    ci = cell_infos(nc)
    data = np.random.default_rng(5).standard_normal((nc, r64.dim * bs)).astype(np.float32)
    want = r32.T_apply_batch(op, data.copy(), bs, ci)
    got = data.copy()
    e.T_apply_batch(got, bs, ci, op=op)
    if r64.dof_transformations_are_permutations:
        assert np.array_equal(got, want)
    else:
        assert np.max(np.abs(got - want)) <= 2e-5 * np.max(np.abs(want))


def test_polyset_float32_vs_reference_float(ref):
    import basix_b200 as bx
    from basix_b200 import _capi

    eps = float(np.finfo(np.float32).eps)
    for cell, degree, nd in ((TRI, 4, 2), (TET, 5, 1), (HEX, 3, 1), (INTERVAL, 6, 2)):
        x32 = random_points(cell, 1025, seed=3).astype(np.float32)
        want = ref.tabulate_polynomial_set_f32(cell, 0, degree, nd, x32)
        exact = ref.tabulate_polynomial_set(cell, 0, degree, nd, x32.astype(np.float64))
        got = np.empty_like(want)
        _capi.check(_capi.lib().bx_polyset_tabulate_f32(cell, 0, degree, nd, x32.ctypes.data, len(x32), got.ctypes.data,
                                                        _capi.BX_MEM_HOST, None))
        for d in range(want.shape[0]):
            scale = np.max(np.abs(exact[d]))
            ref_err = np.max(np.abs(want[d].astype(np.float64) - exact[d]))
            # the float32 polyset kernels also compute in float (their own rounding): no worse than the reference's
            # float path against the float64 truth, and the two agree to the sum of their errors
            our_err = np.max(np.abs(got[d].astype(np.float64) - exact[d]))
            assert our_err <= max(4 * ref_err, 16 * eps * scale), (cell, degree, d, our_err, ref_err, scale)
            assert np.max(np.abs(got[d].astype(np.float64) - want[d])) <= ref_err + our_err + eps * scale


def test_push_forward_broadcast_256_rows(ref):
    """243..256 rows of three doubles need more than 48 KB of dynamic shared memory (opt-in attribute)."""
    r, e = make_pair(ref, N1E, TET, 3, LEG)
    rng = np.random.default_rng(9)
    for rows in (242, 243, 256, 257):
        nc = 301
        J, detJ, K = jacobians(nc, 3, 3, seed=rows)
        U0 = rng.standard_normal((rows, 3))
        want = r.push_forward(np.ascontiguousarray(np.broadcast_to(U0, (nc, rows, 3))), J, detJ, K)
        assert np.array_equal(e.push_forward_broadcast(U0, J, detJ, K), want)


@pytest.mark.parametrize("op", ["T_apply", "Tt_apply_right", "Tinv_apply"])
def test_T_apply_block_size_of_a_whole_element_tensor(ref, op):
    """Block size = ndofs (a caller transforming an element matrix, reference test_dof_transformations.py:45-66) on an
    element whose single cell (216 x 216 doubles) is larger than shared memory: processed in slabs of blocks."""
    r, e = make_pair(ref, P, HEX, 5, GLL)
    n = r.dim
    ci = cell_infos(5)
    data = np.random.default_rng(1).standard_normal((5, n * n))
    want = r.T_apply_batch(op, data.copy(), n, ci)
    got = data.copy()
    e.T_apply_batch(got, n, ci, op=op)
    assert np.array_equal(got, want)  # permutation element: pure moves
    r, e = make_pair(ref, N1E, TET, 3, LEG)
    data = np.random.default_rng(2).standard_normal((5, 45 * 1200))
    want = r.T_apply_batch(op, data.copy(), 1200, cell_infos(5))
    got = data.copy()
    e.T_apply_batch(got, 1200, cell_infos(5), op=op)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))


# ---- macroedge polynomial sets (polyset.cpp:126-1591) ---------------------------------------------------------------
MACRO = [(INTERVAL, 0, 0), (INTERVAL, 1, 0), (INTERVAL, 3, 0), (INTERVAL, 4, 2), (INTERVAL, 2, 1), (QUAD, 0, 1), (QUAD, 1, 0),
         (QUAD, 3, 0), (QUAD, 2, 2), (HEX, 1, 0), (HEX, 2, 0), (HEX, 2, 1), (TRI, 0, 1), (TRI, 1, 0), (TRI, 1, 2), (TRI, 2, 0),
         (TRI, 2, 1), (TRI, 2, 3), (TET, 0, 1), (TET, 1, 0), (TET, 1, 1), (TET, 1, 2)]


@pytest.mark.parametrize("cell,degree,nd", MACRO)
def test_polyset_macroedge(ref, cell, degree, nd):
    """Identical to the reference, including its derivative slabs as they are (the 1-D code multiplies by `-i` for an
    unsigned i, so slabs of order >= 1 hold entries of magnitude 1e19): triangle / tetrahedron bit for bit (same
    expressions), interval / quadrilateral / hexahedron to rounding of pow()."""
    import torch

    import basix_b200 as bx

    x = random_points(cell, 2003, seed=degree + 7 * nd)
    # points on the inner edges and at the vertices of the sub-cells
    x[:4] = {INTERVAL: [[0.5], [0.0], [1.0], [0.25]], TRI: [[0.5, 0.0], [0.25, 0.25], [0.0, 0.5], [0.5, 0.5]],
             QUAD: [[0.5, 0.5], [0.5, 0.1], [0.2, 0.5], [0.0, 1.0]], HEX: [[0.5, 0.5, 0.5], [0.5, 0.2, 0.7], [0, 0, 0], [1, 1, 1]],
             TET: [[0.5, 0, 0], [0.25, 0.25, 0], [0.25, 0.25, 0.25], [0, 0.5, 0.5]]}[cell]
    want = ref.tabulate_polynomial_set(cell, 1, degree, nd, x)
    got = bx.tabulate_polynomial_set(cell, 1, degree, nd, x)
    assert got.shape == want.shape == (bx.polyset_nderivs(cell, nd), bx.polyset_dim(cell, 1, degree), len(x))
    if cell in (TRI, TET):
        assert np.array_equal(got, want)
    else:
        assert np.array_equal(np.isfinite(got), np.isfinite(want))
        for d in range(want.shape[0]):
            for f in range(want.shape[1]):
                w, g = want[d, f], got[d, f]
                assert np.max(np.abs(g - w)) <= 1e-13 * max(np.max(np.abs(w)), 1e-300), (d, f)
    dev = bx.tabulate_polynomial_set(cell, 1, degree, nd, torch.from_numpy(x).cuda())
    assert np.array_equal(dev.cpu().numpy(), got)


def test_polyset_macroedge_orthonormal_and_limits(ref):
    """Reference test_orthonormal_basis.py:120-146: the sets are orthonormal (piecewise quadrature: the Gauss-Jacobi rule
    of each sub-interval / sub-cell)."""
    import basix_b200 as bx

    # interval, degree 3: integrate over [0, 1/2] and [1/2, 1] separately
    g, w = np.polynomial.legendre.leggauss(8)
    xs = np.concatenate([(g + 1) / 4, (g + 1) / 4 + 0.5])[:, None]
    ws = np.concatenate([w / 4, w / 4])
    for degree in range(0, 5):
        B = bx.tabulate_polynomial_set(INTERVAL, 1, degree, 0, xs)[0]
        assert np.allclose((B * ws) @ B.T, np.eye(2 * degree + 1), atol=1e-12)
    with pytest.raises(RuntimeError, match="Only degree 0 to 2"):
        bx.tabulate_polynomial_set(TRI, 1, 3, 0, np.zeros((1, 2)))
    with pytest.raises(RuntimeError, match="Only degree 0 and 1"):
        bx.tabulate_polynomial_set(TET, 1, 2, 0, np.zeros((1, 3)))
