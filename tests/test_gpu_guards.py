"""Out-of-bounds and determinism checks of our own (compute-sanitizer is closed on this GPU pool).

Every kernel family is called through the C ABI on device buffers that sit between two guard zones filled with a byte
pattern: the guards must be untouched afterwards (no stray store before the first or past the last entry, for ragged
sizes that end inside a tile), read-only inputs must be unchanged, and a second run on the same inputs must give the
same bits (the kernels with warp-private staging, TMA bulk copies and mbarrier hand-offs have no data race that shows).
"""

import ctypes

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

GUARD = 4096  # bytes on either side
PATTERN = 0xA5


def _lib():
    from basix_b200 import _capi

    return _capi, _capi.lib()


class Guarded:
    """A device buffer of `count` elements of `dtype` between two guard zones."""

    def __init__(self, count, dtype, fill=None):
        import torch

        self.itemsize = np.dtype(dtype).itemsize
        self.nbytes = int(count) * self.itemsize
        self.raw = torch.full((GUARD + self.nbytes + GUARD,), PATTERN, dtype=torch.uint8, device="cuda")
        self.body = self.raw[GUARD:GUARD + self.nbytes]
        tdt = {np.float64: torch.float64, np.float32: torch.float32, np.int32: torch.int32, np.uint32: torch.int32}[dtype]
        self.view = self.body.view(tdt)
        if fill is not None:
            src = torch.from_numpy(np.ascontiguousarray(fill).view(np.int32 if dtype is np.uint32 else dtype).reshape(-1))
            self.view.copy_(src.cuda())

    @property
    def ptr(self):
        return ctypes.c_void_p(self.body.data_ptr())

    def intact(self):
        import torch

        torch.cuda.synchronize()
        return bool((self.raw[:GUARD] == PATTERN).all()) and bool((self.raw[GUARD + self.nbytes:] == PATTERN).all())

    def snapshot(self):
        import torch

        torch.cuda.synchronize()
        return self.body.clone()


def _run_twice(call, outs, ins):
    """call() twice; guards intact, inputs unchanged, outputs bit-identical between the runs."""
    import torch

    before = [b.snapshot() for b in ins]
    call()
    first = [b.snapshot() for b in outs]
    for b in outs:
        if b not in ins:
            b.body.fill_(PATTERN)
    for b, s in zip(ins, before):
        if b in outs:  # in-place: restore the input
            b.body.copy_(s)
    call()
    second = [b.snapshot() for b in outs]
    for b in list(outs) + list(ins):
        assert b.intact(), "a guard zone was written"
    for b, s in zip(ins, before):
        if b not in outs:
            assert torch.equal(b.snapshot(), s), "a read-only input was modified"
    for a, b in zip(first, second):
        assert torch.equal(a, b), "two runs on the same inputs differ"
    return first


def _points(cell, n, seed=0):
    rng = np.random.default_rng(seed)
    tdim = {1: 1, 2: 2, 3: 3, 4: 2, 5: 3}[cell]
    x = rng.random((n, tdim))
    if cell in (2, 3):
        x /= tdim + 0.5  # inside the simplex
    return np.ascontiguousarray(x)


# (family, cell, degree, variant, nd, npts): one case per tabulate kernel, sizes that end inside a tile
TABULATE = [
    (1, 3, 5, 2, 2, 1000 + 13),   # derivative-operator kernel, 16-point chunks
    (1, 2, 1, 2, 1, 4096 + 37),   # register kernel, 128-point tiles
    (1, 2, 2, 2, 2, 257),         # register kernel, other shape
    (1, 5, 4, 2, 1, 2048 + 31),   # staged tensor-product kernel
    (1, 4, 6, 2, 2, 333),         # staged tensor-product, quadrilateral
    (1, 5, 5, 2, 1, 95),          # direct tensor-product kernel
    (1, 3, 6, 2, 2, 211),         # dense fused kernel (operators too large for shared memory)
    (3, 3, 3, 11, 1, 129),        # vector valued, derivative operator
    (2, 3, 4, 11, 0, 77),         # RT4
]


@pytest.mark.parametrize("family,cell,degree,variant,nd,npts", TABULATE)
@pytest.mark.parametrize("dtype", [np.float64, np.float32])
def test_tabulate_stays_inside_its_table(family, cell, degree, variant, nd, npts, dtype):
    import basix_b200 as bx

    capi, L = _lib()
    e = bx.create_element(family, cell, degree, variant)
    shape = e.tabulate_shape(nd, npts)
    x = Guarded(npts * {1: 1, 2: 2, 3: 3, 4: 2, 5: 3}[cell], dtype, _points(cell, npts).astype(dtype))
    out = Guarded(int(np.prod(shape)), dtype)
    fn = L.bx_tabulate_f32 if dtype is np.float32 else L.bx_tabulate_f64
    tdim = {1: 1, 2: 2, 3: 3, 4: 2, 5: 3}[cell]

    def call():
        assert fn(e._h, nd, x.ptr, npts, tdim, out.ptr, capi.BX_MEM_DEVICE, None) == 0, L.bx_last_error()

    (res,) = _run_twice(call, [out], [x])
    import torch

    assert not bool(torch.isnan(res.view(torch.float32 if dtype is np.float32 else torch.float64)).any())


def _jac(nc, gdim, tdim, seed):
    rng = np.random.default_rng(seed)
    J = np.zeros((nc, gdim, tdim))
    J[:, :tdim, :tdim] = np.eye(tdim)
    J += 0.2 * rng.standard_normal((nc, gdim, tdim))
    if gdim == tdim:
        return J, np.linalg.det(J), np.linalg.inv(J)
    JtJ = np.einsum("cji,cjk->cik", J, J)
    return J, np.sqrt(np.linalg.det(JtJ)), np.einsum("cij,ckj->cik", np.linalg.inv(JtJ), J)


@pytest.mark.parametrize("family,cell,degree,gdim", [(3, 3, 3, 3), (2, 3, 4, 3), (3, 2, 2, 3), (2, 2, 2, 2)])
@pytest.mark.parametrize("rows,nc", [(1, 1000 + 7), (45, 333), (33, 65)])
def test_maps_stay_inside_their_arrays(family, cell, degree, gdim, rows, nc):
    import basix_b200 as bx

    capi, L = _lib()
    e = bx.create_element(family, cell, degree, 11)
    tdim = 3 if cell == 3 else 2
    Jn, dn, Kn = _jac(nc, gdim, tdim, rows)
    U = Guarded(nc * rows * tdim, np.float64, np.random.default_rng(1).standard_normal(nc * rows * tdim))
    U1 = Guarded(rows * tdim, np.float64, np.random.default_rng(2).standard_normal(rows * tdim))
    J, d, K = Guarded(Jn.size, np.float64, Jn), Guarded(dn.size, np.float64, dn), Guarded(Kn.size, np.float64, Kn)
    out = Guarded(nc * rows * gdim, np.float64)
    back = Guarded(nc * rows * tdim, np.float64)
    M = capi.BX_MEM_DEVICE

    def push():
        assert L.bx_push_forward_f64(e._h, U.ptr, J.ptr, d.ptr, K.ptr, nc, rows, tdim, gdim, tdim, out.ptr, M, None) == 0

    def bcast():
        assert L.bx_push_forward_broadcast_f64(e._h, U1.ptr, J.ptr, d.ptr, K.ptr, nc, rows, tdim, gdim, tdim, out.ptr, M,
                                               None) == 0

    def pull():
        assert L.bx_pull_back_f64(e._h, out.ptr, J.ptr, d.ptr, K.ptr, nc, rows, gdim, gdim, tdim, back.ptr, M, None) == 0

    _run_twice(bcast, [out], [U1, J, d, K])
    _run_twice(push, [out], [U, J, d, K])
    _run_twice(pull, [back], [out, J, d, K])


@pytest.mark.parametrize("family,cell,degree,variant,n", [(2, 3, 4, 11, 1), (2, 3, 4, 11, 3), (3, 3, 3, 11, 1), (1, 3, 5, 2, 1),
                                                         (1, 5, 3, 2, 2), (3, 2, 3, 11, 1)])
@pytest.mark.parametrize("nc", [1, 63, 64 * 5 + 17])
def test_dof_transformations_stay_inside_their_cells(family, cell, degree, variant, n, nc):
    import basix_b200 as bx

    capi, L = _lib()
    e = bx.create_element(family, cell, degree, variant)
    rng = np.random.default_rng(nc)
    ci = Guarded(nc, np.uint32, rng.integers(0, 2 ** 30, nc).astype(np.uint32))
    data = Guarded(nc * e.dim * n, np.float64, rng.standard_normal(nc * e.dim * n))
    M = capi.BX_MEM_DEVICE
    for op in range(8):
        def call(op=op):
            assert L.bx_T_apply_f64(e._h, op, data.ptr, nc, e.dim * n, n, ci.ptr, M, None) == 0, L.bx_last_error()

        _run_twice(call, [data], [data, ci])
    if e.dof_transformations_are_permutations:
        d = Guarded(nc * e.dim, np.int32, np.tile(np.arange(e.dim, dtype=np.int32), nc))
        for inv in (0, 1):
            def perm(inv=inv):
                assert L.bx_permute_i32(e._h, inv, d.ptr, nc, ci.ptr, M, None) == 0

            _run_twice(perm, [d], [d, ci])


@pytest.mark.parametrize("family,degree,npts,nc", [(3, 3, 1, 1000 + 3), (2, 4, 1, 97), (3, 2, 4, 65), (2, 2, 3, 33)])
def test_geometry_and_physical_basis_stay_inside_their_arrays(family, degree, npts, nc):
    import basix_b200 as bx

    capi, L = _lib()
    e = bx.create_element(family, 3, degree, 11)
    rng = np.random.default_rng(degree)
    verts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]])
    cn = verts[None] + 0.1 * rng.standard_normal((nc, 4, 3))
    coords = Guarded(cn.size, np.float64, cn)
    ci = Guarded(nc, np.uint32, rng.integers(0, 2 ** 18, nc).astype(np.uint32))
    X = Guarded(npts * 3, np.float64, _points(3, npts, seed=3))
    J, d, K = Guarded(nc * 9, np.float64), Guarded(nc, np.float64), Guarded(nc * 9, np.float64)
    out = Guarded(nc * npts * e.dim * 3, np.float64)
    M = capi.BX_MEM_DEVICE

    def geom():
        assert L.bx_geometry_f64(coords.ptr, None, nc, 4, 3, 3, 1, J.ptr, d.ptr, K.ptr, M, None) == 0, L.bx_last_error()

    def fused():
        assert L.bx_physical_basis_f64(e._h, 0, X.ptr, npts, 3, coords.ptr, None, None, None, ci.ptr, nc, 3, out.ptr, M,
                                       None) == 0, L.bx_last_error()

    def fused_jk():
        assert L.bx_physical_basis_f64(e._h, 0, X.ptr, npts, 3, None, J.ptr, d.ptr, K.ptr, ci.ptr, nc, 3, out.ptr, M,
                                       None) == 0, L.bx_last_error()

    _run_twice(geom, [J, d, K], [coords])
    (a,) = _run_twice(fused, [out], [coords, ci, X])
    (b,) = _run_twice(fused_jk, [out], [J, d, K, ci, X])
    import torch

    assert torch.equal(a, b)  # geometry on chip == geometry from bx_geometry


@pytest.mark.parametrize("family,cell,degree,variant,nc", [(3, 3, 3, 11, 1000 + 9), (2, 3, 2, 11, 77), (1, 3, 4, 2, 130)])
@pytest.mark.parametrize("layout", [0, 1])
def test_interpolate_stays_inside_its_arrays(family, cell, degree, variant, nc, layout):
    import basix_b200 as bx

    capi, L = _lib()
    e = bx.create_element(family, cell, degree, variant)
    ndofs, ncols = e.interpolation_matrix.shape
    vals = Guarded(nc * ncols, np.float64, np.random.default_rng(5).standard_normal(nc * ncols))
    coef = Guarded(nc * ndofs, np.float64)

    def call():
        assert L.bx_interpolate_f64(e._h, layout, vals.ptr, nc, coef.ptr, capi.BX_MEM_DEVICE, None) == 0, L.bx_last_error()

    _run_twice(call, [coef], [vals])


@pytest.mark.parametrize("family,degree,nc", [(3, 3, 1000 + 9), (2, 2, 77), (3, 1, 17)])
def test_interpolate_pulled_stays_inside_its_arrays(family, degree, nc):
    import basix_b200 as bx

    capi, L = _lib()
    e = bx.create_element(family, 3, degree, 11)
    npts = e.points.shape[0]
    Jn, dn, Kn = _jac(nc, 3, 3, degree)
    u = Guarded(nc * npts * 3, np.float64, np.random.default_rng(6).standard_normal(nc * npts * 3))
    J, d, K = Guarded(Jn.size, np.float64, Jn), Guarded(dn.size, np.float64, dn), Guarded(Kn.size, np.float64, Kn)
    coef = Guarded(nc * e.dim, np.float64)

    def call():
        assert L.bx_interpolate_pulled_f64(e._h, u.ptr, J.ptr, d.ptr, K.ptr, nc, 3, coef.ptr, capi.BX_MEM_DEVICE, None) == 0, \
            L.bx_last_error()

    _run_twice(call, [coef], [u, J, d, K])
