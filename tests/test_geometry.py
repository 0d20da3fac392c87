"""Cell geometry (bx_geometry_*) and the fused tabulate-once -> T_apply -> push_forward pass (bx_physical_basis_* /
bx_push_forward_fused_*), SURVEY.md 8f rank 3.

Oracle: the reference has the affine case only (cell::entity_jacobians, cell.cpp:662-689); oracle/restate.py::geometry
restates the library's documented arithmetic operation for operation and is pinned here (CPU) to the reference's
entity_jacobians bit for bit and to numpy.linalg det / inv / pinv.  The fused pass is compared with the reference's own
three calls (T_apply on the replicated reference values, then push_forward) to 1e-12, and with the library's own
three-call sequence bit for bit."""

import numpy as np
import pytest

from conftest import make_pair, random_points

P, RT, N1E = 1, 2, 3
INTERVAL, TRI, TET, QUAD, HEX = 1, 2, 3, 4, 5
GLL, EQ, LEG = 2, 1, 11
SHAPES = [(1, 1), (2, 2), (3, 3), (2, 1), (3, 1), (3, 2)]


def p1_derivatives(tdim, dtype=np.float64):
    """Derivative slabs of the P1 simplex coordinate element, exact: (tdim, 1, tdim + 1)."""
    d = np.zeros((tdim, 1, tdim + 1), dtype=dtype)
    d[:, 0, 0] = -1
    for t in range(tdim):
        d[t, 0, 1 + t] = 1
    return d


# ---- the restatement is pinned (CPU) -------------------------------------------------------------------------------
def test_restated_geometry_matches_reference_entity_jacobians(ref):
    from oracle import restate as R

    for cell, tdim in ((TRI, 2), (TET, 3), (QUAD, 2), (HEX, 3), (6, 3), (7, 3)):
        for e_dim in range(1, tdim):
            want = ref.entity_jacobians(cell, e_dim)
            for i in range(want.shape[0]):
                x = ref.sub_entity_geometry(cell, e_dim, i)
                J, _, _ = R.geometry(x[None, :e_dim + 1, :])
                assert np.array_equal(J[0], want[i])
                # the general sum with the exact P1 derivatives is the same difference
                J2, _, _ = R.geometry(x[None, :e_dim + 1, :], p1_derivatives(e_dim))
                assert np.array_equal(J2[0, 0], want[i])


@pytest.mark.parametrize("gdim,tdim", SHAPES)
def test_restated_geometry_against_numpy_linalg(gdim, tdim):
    from oracle import restate as R

    x = np.random.default_rng(gdim * 4 + tdim).standard_normal((200, tdim + 1, gdim))
    J, det, K = R.geometry(x)
    if gdim == tdim:
        assert np.allclose(det, np.linalg.det(J), rtol=1e-12, atol=1e-13)
        assert np.allclose(np.einsum("cij,cjk->cik", K, J), np.eye(tdim)[None], atol=1e-9)
    else:
        assert np.allclose(K, np.linalg.pinv(J), rtol=1e-9, atol=1e-10)
        assert np.allclose(det, np.sqrt(np.linalg.det(np.einsum("cgi,cgj->cij", J, J))), rtol=1e-12)


# ---- bx_geometry ------------------------------------------------------------------------------------------------------
@pytest.mark.gpu
@pytest.mark.parametrize("gdim,tdim", SHAPES)
def test_geometry_affine_bit_exact(ref, gdim, tdim):
    import torch

    import basix_b200 as bx
    from oracle import restate as R

    for nc in (1, 33, 4099):
        x = np.random.default_rng(nc).standard_normal((nc, tdim + 1, gdim))
        want = R.geometry(x)
        got = bx.geometry(x)
        for g, w in zip(got, want):
            assert g.shape == w.shape and np.array_equal(g, w)
        dev = bx.geometry(torch.from_numpy(x).cuda())
        for g, w in zip(dev, want):
            assert g.is_cuda and np.array_equal(g.cpu().numpy(), w)
        # the general sum with the exact derivatives of the P1 coordinate element gives the same bits
        gen = bx.geometry(x, p1_derivatives(tdim))
        for g, w in zip(gen, want):
            assert np.array_equal(g.reshape(w.shape), w)
        x32 = x.astype(np.float32)
        for g, w in zip(bx.geometry(x32), R.geometry(x32)):
            assert g.dtype == np.float32 and np.array_equal(g, w)
    only_K = bx.geometry(x, want=("K",))
    assert only_K[0] is None and only_K[1] is None and np.array_equal(only_K[2], want[2])


@pytest.mark.gpu
def test_geometry_reference_sub_entities(ref):
    """The kernel on the reference cell's own sub-entities reproduces cell::entity_jacobians bit for bit."""
    import basix_b200 as bx

    for cell, tdim in ((TRI, 2), (TET, 3), (HEX, 3), (6, 3)):
        for e_dim in range(1, tdim):
            want = ref.entity_jacobians(cell, e_dim)
            x = np.stack([ref.sub_entity_geometry(cell, e_dim, i)[:e_dim + 1] for i in range(want.shape[0])])
            J, detJ, K = bx.geometry(x)
            assert np.array_equal(J, want)
            assert np.allclose(np.einsum("cij,cjk->cik", K, J), np.eye(e_dim)[None], atol=1e-14)


@pytest.mark.gpu
@pytest.mark.parametrize("cell,degree,gdim", [(TRI, 2, 2), (TRI, 2, 3), (QUAD, 1, 2), (QUAD, 2, 3), (TET, 2, 3), (HEX, 1, 3),
                                              (HEX, 2, 3), (INTERVAL, 2, 1), (INTERVAL, 3, 3)])
def test_geometry_higher_order_cells(ref, cell, degree, gdim):
    """Non-affine cells: J varies from point to point; derivatives from the coordinate element's own tabulate."""
    import basix_b200 as bx
    from oracle import restate as R

    r, e = make_pair(ref, P, cell, degree, GLL)
    tdim = r.tdim
    X = random_points(cell, 13, seed=degree)
    dphi = np.ascontiguousarray(r.tabulate(1, X)[1:, :, :, 0])
    nodes = r.points  # the reference cell's own nodes, perturbed
    rng = np.random.default_rng(cell)
    nc = 517
    coords = np.zeros((nc, r.dim, gdim))
    coords[:, :, :tdim] = nodes[None]
    coords += 0.05 * rng.standard_normal(coords.shape)
    want = R.geometry(coords, dphi)
    got = bx.geometry(coords, dphi)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
    J, detJ, K = got
    assert np.allclose(np.einsum("cpij,cpjk->cpik", K, J), np.eye(tdim)[None, None], atol=1e-9)
    # the library's own tabulate of the coordinate element feeds the same call on the device
    import torch

    dphi_dev = e.tabulate(1, torch.from_numpy(X).cuda())[1:, :, :, 0].contiguous()
    Jd, dd, Kd = bx.geometry(torch.from_numpy(coords).cuda(), dphi_dev)
    assert np.allclose(Jd.cpu().numpy(), J, rtol=1e-12, atol=1e-13) and np.allclose(Kd.cpu().numpy(), K, rtol=1e-10, atol=1e-11)


@pytest.mark.gpu
@pytest.mark.parametrize("npts", [1, 8, 33, 64])
@pytest.mark.parametrize("nc", [1, 101])
def test_geometry_general_form_tile_edges(ref, npts, nc):
    """The warp form of the general kernel: point counts below, at and above the 32-item tile, ragged totals, outputs
    asked for one at a time -- bit for bit against the restatement."""
    import basix_b200 as bx
    from oracle import restate as R

    r, _ = make_pair(ref, P, HEX, 1, GLL)
    X = random_points(HEX, npts, seed=npts)
    dphi = np.ascontiguousarray(r.tabulate(1, X)[1:, :, :, 0])
    coords = r.points[None] + 0.05 * np.random.default_rng(nc).standard_normal((nc, 8, 3))
    want = R.geometry(coords, dphi)
    got = bx.geometry(coords, dphi)
    for g, w in zip(got, want):
        assert g.shape == w.shape and np.array_equal(g, w)
    for k, name in enumerate(("J", "detJ", "K")):
        one = bx.geometry(coords, dphi, want=(name,))
        assert np.array_equal(one[k], want[k]) and all(one[j] is None for j in range(3) if j != k)


@pytest.mark.gpu
def test_geometry_errors():
    import basix_b200 as bx

    with pytest.raises(RuntimeError, match="affine simplices"):
        bx.geometry(np.zeros((3, 5, 3)), tdim=3)
    with pytest.raises(RuntimeError, match="tdim <= gdim"):
        bx.geometry(np.zeros((3, 4, 2)))
    assert bx.geometry(np.zeros((0, 4, 3)))[0].shape == (0, 3, 3)


# ---- fused physical basis ----------------------------------------------------------------------------------------------
def cell_infos(nc, seed=11):
    rng = np.random.default_rng(seed)
    ci = rng.integers(0, 2 ** 30, nc).astype(np.uint32)
    ci[:4] = [0, 1, 0x2ABCD, 0x3FFFFFFF]
    return ci


FUSED = [(N1E, TET, 3, LEG), (RT, TET, 4, LEG), (RT, TRI, 2, LEG), (N1E, TRI, 2, LEG), (P, TET, 3, GLL), (P, TRI, 4, GLL),
         (N1E, TET, 1, LEG), (P, TET, 5, GLL)]


@pytest.mark.gpu
@pytest.mark.parametrize("family,cell,degree,variant", FUSED)
@pytest.mark.parametrize("npts", [1, 7])
def test_physical_basis_against_three_calls(ref, family, cell, degree, variant, npts):
    import torch

    import basix_b200 as bx
    from oracle import restate as R

    r, e = make_pair(ref, family, cell, degree, variant)
    tdim, nd, vs = r.tdim, r.dim, r.value_size
    nc = 1237
    rng = np.random.default_rng(degree + npts)
    ref_x = np.zeros((tdim + 1, tdim))
    ref_x[1:] = np.eye(tdim)
    coords = ref_x[None] + 0.2 * rng.standard_normal((nc, tdim + 1, tdim))
    ci = cell_infos(nc)
    X = random_points(cell, npts, seed=3)
    got = e.physical_basis(X, coords=coords, cell_info=ci)
    out_vs = vs if r.map_type == 0 else tdim
    assert got.shape == (nc, npts, nd, out_vs)

    # (1) the reference's own three calls
    U = r.tabulate(0, X)[0]  # (npts, ndofs, vs)
    rep = np.ascontiguousarray(np.broadcast_to(U.reshape(1, npts, nd * vs), (nc, npts, nd * vs))).reshape(nc * npts, nd * vs)
    r.T_apply_batch("T_apply", rep, vs, np.repeat(ci, npts))
    J, detJ, K = R.geometry(coords)
    want = r.push_forward(rep.reshape(nc, npts * nd, vs), J, detJ, K).reshape(nc, npts, nd, out_vs)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))

    # (2) the library's own three calls: identical bits
    Ub = e.tabulate(0, X)[0]
    repb = np.ascontiguousarray(np.broadcast_to(Ub.reshape(1, npts, nd * vs), (nc, npts, nd * vs))).reshape(nc * npts, nd * vs)
    e.T_apply_batch(repb, vs, np.repeat(ci, npts))
    Jb, db, Kb = bx.geometry(coords)
    seq = e.push_forward(repb.reshape(nc, npts * nd, vs), Jb, db, Kb).reshape(nc, npts, nd, out_vs)
    assert np.array_equal(got, seq)

    # device pointers, reference values given, geometry given instead of coordinates
    dev = e.physical_basis(torch.from_numpy(X).cuda(), coords=torch.from_numpy(coords).cuda(),
                           cell_info=torch.from_numpy(ci.view(np.int32)).cuda())
    assert dev.is_cuda and np.array_equal(dev.cpu().numpy(), got)
    assert np.array_equal(e.physical_basis(None, U=Ub, coords=coords, cell_info=ci), got)
    if r.map_type != 0:
        assert np.array_equal(e.physical_basis(X, J=Jb, detJ=db, K=Kb, cell_info=ci), got)
    # no cell_info: plain broadcast push
    plain = e.physical_basis(X, coords=coords)
    want_plain = e.push_forward_broadcast(Ub.reshape(npts * nd, vs), Jb, db, Kb).reshape(nc, npts, nd, out_vs) \
        if r.map_type != 0 else np.broadcast_to(Ub, (nc, npts, nd, vs))
    assert np.array_equal(plain, want_plain)


@pytest.mark.gpu
@pytest.mark.parametrize("family,cell,degree,variant", [(N1E, TET, 3, LEG), (RT, TET, 2, LEG), (RT, TRI, 2, LEG), (N1E, TRI, 1, LEG)])
@pytest.mark.parametrize("npts,nc", [(2, 5), (5, 517), (13, 100)])
def test_physical_basis_per_point_geometry(ref, family, cell, degree, variant, npts, nc):
    """Non-affine cells (curved P2 simplices): the Jacobian varies from point to point.  bx.geometry(coords, dphi) feeds
    the fused pass, which must equal -- bit for bit -- T_apply on the replicated reference values followed by
    push_forward with one Jacobian per (cell, point); and the reference's own calls to 1e-12.  Host and device pointers."""
    import torch

    import basix_b200 as bx

    r, e = make_pair(ref, family, cell, degree, variant)
    rg, _ = make_pair(ref, P, cell, 2, GLL)  # the P2 coordinate element
    tdim, nd, vs = r.tdim, r.dim, r.value_size
    rng = np.random.default_rng(npts + nc)
    X = random_points(cell, npts, seed=9)
    dphi = np.ascontiguousarray(rg.tabulate(1, X)[1:, :, :, 0])
    coords = rg.points[None] + 0.04 * rng.standard_normal((nc, rg.dim, tdim))
    J, detJ, K = bx.geometry(coords, dphi)  # (nc, npts, ...)
    ci = cell_infos(nc)
    got = e.physical_basis(X, J=J, detJ=detJ, K=K, cell_info=ci)
    assert got.shape == (nc, npts, nd, tdim)
    # the library's own separate calls
    Ub = e.tabulate(0, X)[0]
    rep = np.ascontiguousarray(np.broadcast_to(Ub.reshape(1, npts, nd * vs), (nc, npts, nd * vs))).reshape(nc * npts, nd * vs)
    e.T_apply_batch(rep, vs, np.repeat(ci, npts))
    two = e.push_forward(rep.reshape(nc * npts, nd, vs), J.reshape(nc * npts, tdim, tdim), detJ.reshape(-1),
                         K.reshape(nc * npts, tdim, tdim)).reshape(nc, npts, nd, tdim)
    assert np.array_equal(got, two)
    # the reference's
    Ur = r.tabulate(0, X)[0]
    repr_ = np.ascontiguousarray(np.broadcast_to(Ur.reshape(1, npts, nd * vs), (nc, npts, nd * vs))).reshape(nc * npts, nd * vs)
    r.T_apply_batch("T_apply", repr_, vs, np.repeat(ci, npts))
    want = r.push_forward(repr_.reshape(nc * npts, nd, vs), J.reshape(nc * npts, tdim, tdim), detJ.reshape(-1),
                          K.reshape(nc * npts, tdim, tdim)).reshape(nc, npts, nd, tdim)
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    # device pointers, and the table given instead of the points
    dev = [torch.from_numpy(a).cuda() for a in (J, detJ, K)]
    got_dev = e.physical_basis(None, U=torch.from_numpy(Ub).cuda(), J=dev[0], detJ=dev[1], K=dev[2],
                               cell_info=torch.from_numpy(ci.view(np.int32)).cuda())
    assert np.array_equal(got_dev.cpu().numpy(), got)
    got32 = e.physical_basis(X.astype(np.float32), J=J.astype(np.float32), detJ=detJ.astype(np.float32), K=K.astype(np.float32),
                             cell_info=ci)
    assert got32.dtype == np.float32 and np.allclose(got32, want, rtol=3e-4, atol=3e-4 * np.max(np.abs(want)))


@pytest.mark.gpu
def test_physical_basis_other_operators_manifold_and_float32(ref):
    import basix_b200 as bx

    # Tt_inv_apply (what a caller applies to functionals), triangle cells embedded in 3-D
    r, e = make_pair(ref, N1E, TRI, 2, LEG)
    nc, npts = 301, 3
    rng = np.random.default_rng(5)
    coords = rng.standard_normal((nc, 3, 3))
    ci = cell_infos(nc, seed=4)
    X = random_points(TRI, npts)
    from oracle import restate as R

    J, detJ, K = R.geometry(coords)
    for op in ("Tt_inv_apply", "Tinv_apply", "Tt_apply"):
        got = e.physical_basis(X, coords=coords, cell_info=ci, op=op)
        U = r.tabulate(0, X)[0]
        rep = np.ascontiguousarray(np.broadcast_to(U.reshape(1, npts, -1), (nc, npts, r.dim * 2))).reshape(nc * npts, -1)
        r.T_apply_batch(op, rep, 2, np.repeat(ci, npts))
        want = r.push_forward(rep.reshape(nc, npts * r.dim, 2), J, detJ, K).reshape(nc, npts, r.dim, 3)
        assert got.shape == want.shape and np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    # float32 arrays
    r, e = make_pair(ref, RT, TET, 2, LEG)
    coords = (np.eye(4, 3, k=-1)[None] + 0.2 * rng.standard_normal((nc, 4, 3))).astype(np.float32)
    X32 = random_points(TET, 4).astype(np.float32)
    got = e.physical_basis(X32, coords=coords, cell_info=ci)
    want = e.physical_basis(X32.astype(np.float64), coords=coords.astype(np.float64), cell_info=ci)
    assert got.dtype == np.float32 and np.max(np.abs(got - want)) <= 2e-5 * np.max(np.abs(want))


@pytest.mark.gpu
def test_physical_basis_errors(ref):
    r, e = make_pair(ref, 7, TET, 1, 0)  # Regge: double covariant Piola
    with pytest.raises(RuntimeError, match="double Piola"):
        e.physical_basis(np.zeros((1, 3)), coords=np.zeros((2, 4, 3)))
    r, e = make_pair(ref, N1E, TET, 3, LEG)
    with pytest.raises(RuntimeError, match="does not match element dim"):
        e.physical_basis(np.zeros((1, 2)), coords=np.zeros((2, 4, 3)))
    with pytest.raises(RuntimeError, match="needs K"):
        e.physical_basis(np.zeros((1, 3)), J=np.zeros((2, 3, 3)), detJ=np.zeros(2))
    assert e.physical_basis(np.zeros((2, 3)), coords=np.zeros((0, 4, 3))).shape == (0, 2, 45, 3)
