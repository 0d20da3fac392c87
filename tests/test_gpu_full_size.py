"""Parity at the BASELINE.json sizes, through size-independent properties plus a random sample checked against the
compiled reference (the oracle is far too slow for the whole array):

* tabulate: a Lagrange basis sums to 1 at every point and every derivative slab sums to 0
  (reference test_lagrange.py:857-864);
* push_forward -> pull_back is the identity (reference test_mappings.py:24-97);
* T_apply -> Tinv_apply and permute -> permute_inv are the identity (reference test_dof_transformations.py:45-66),
  and a permuted row is still a permutation (a checksum of checksums over all cells).
"""

import numpy as np
import pytest

from conftest import make_pair

pytestmark = pytest.mark.gpu

P, RT, N1E = 1, 2, 3
TET, HEX = 3, 5
GLL, LEG = 2, 11


def _free():
    import torch

    torch.cuda.empty_cache()


def _sample(n, k, seed):
    return np.sort(np.random.default_rng(seed).choice(n, size=k, replace=False))


def _tet_points(n, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    s, _ = torch.sort(torch.rand(n, 3, generator=g, device="cuda", dtype=torch.float64), dim=1)
    return torch.stack([s[:, 0], s[:, 1] - s[:, 0], s[:, 2] - s[:, 1]], dim=1).contiguous()


def test_tabulate_p5_tet_10m_points(ref):
    """BASELINE config 2: P5 tetrahedron, nd = 2, 10 M points (44.8 GB table)."""
    import torch

    r, e = make_pair(ref, P, TET, 5, GLL)
    n = 10_000_000
    x = _tet_points(n, 42)
    t = e.tabulate(2, x)
    assert tuple(t.shape) == (10, n, 56, 1)
    s = t.sum(dim=2)[..., 0]  # [10][n]
    assert float((s[0] - 1).abs().max()) < 1e-11
    scale = max(float(t.max()), -float(t.min()))
    assert float(s[1:].abs().max()) < 1e-11 * scale
    idx = _sample(n, 2000, 1)
    it = torch.from_numpy(idx).cuda()
    want = r.tabulate(2, x[it].cpu().numpy())
    got = t[:, it].cpu().numpy()
    for d in range(10):
        assert np.max(np.abs(got[d] - want[d])) <= 1e-12 * np.max(np.abs(want[d]))
    del t, s
    _free()


def test_tabulate_q4_hex_10m_points(ref):
    """BASELINE config 3 streams 50 M points as 5 launches of 10 M through one resident table; this is one launch."""
    import torch

    r, e = make_pair(ref, P, HEX, 4, GLL)
    n = 10_000_000
    x = torch.rand(n, 3, generator=torch.Generator(device="cuda").manual_seed(42), device="cuda", dtype=torch.float64)
    t = e.tabulate(1, x)
    assert tuple(t.shape) == (4, n, 125, 1)
    s = t.sum(dim=2)[..., 0]
    assert float((s[0] - 1).abs().max()) < 1e-11
    assert float(s[1:].abs().max()) < 1e-11 * max(float(t.max()), -float(t.min()))
    idx = _sample(n, 2000, 2)
    it = torch.from_numpy(idx).cuda()
    want = r.tabulate(1, x[it].cpu().numpy())
    got = t[:, it].cpu().numpy()
    for d in range(4):
        assert np.max(np.abs(got[d] - want[d])) <= 1e-12 * np.max(np.abs(want[d]))
    del t, s
    _free()


def test_push_pull_round_trip_20m_cells(ref):
    """BASELINE config 4: N1curl degree 3 tetrahedron, covariant Piola over 20 M cells."""
    import torch

    r, e = make_pair(ref, N1E, TET, 3, LEG)
    nc = 20_000_000
    g = torch.Generator(device="cuda").manual_seed(7)
    J = torch.eye(3, dtype=torch.float64, device="cuda")[None] + 0.3 * torch.randn(nc, 3, 3, generator=g, device="cuda",
                                                                                 dtype=torch.float64)

    def det_and_inverse(J):
        a, b, c, d, e_, f, g_, h, i = (J[:, r_, c_] for r_ in range(3) for c_ in range(3))
        co = torch.stack([e_ * i - f * h, c * h - b * i, b * f - c * e_,
                          f * g_ - d * i, a * i - c * g_, c * d - a * f,
                          d * h - e_ * g_, b * g_ - a * h, a * e_ - b * d], dim=1)
        det = a * co[:, 0] + b * co[:, 3] + c * co[:, 6]
        return det.contiguous(), (co / det[:, None]).reshape(-1, 3, 3).contiguous()

    detJ, K = det_and_inverse(J)
    bad = detJ.abs() < 0.1
    J[bad] = torch.eye(3, dtype=torch.float64, device="cuda")
    detJ, K = det_and_inverse(J)
    U = torch.randn(nc, 1, 3, generator=g, device="cuda", dtype=torch.float64)
    u = e.push_forward(U, J, detJ, K)
    back = e.pull_back(u, J, detJ, K)
    # K was inverted in floating point: the round trip is exact to the conditioning of J
    assert float(((back - U).abs().amax(dim=(1, 2)) / (1 + U.abs().amax(dim=(1, 2)))).max()) < 1e-9
    idx = torch.from_numpy(_sample(nc, 5000, 3)).cuda()
    a = [v[idx].cpu().numpy() for v in (U, J, detJ, K)]
    assert np.array_equal(u[idx].cpu().numpy(), r.push_forward(*a))  # bit-exact
    assert np.array_equal(back[idx].cpu().numpy(), r.pull_back(u[idx].cpu().numpy(), *a[1:]))
    del J, K, U, u, back
    _free()


def _cell_info(nc, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    ci = torch.randint(0, 1 << 30, (nc,), generator=g, device="cuda", dtype=torch.int64)
    return ci.to(torch.uint32), ci  # the uint32 array the library takes, and an indexable copy


def test_T_apply_round_trip_50m_cells(ref):
    """BASELINE config 5 (i): RT degree 4 tetrahedron, T_apply over 50 M cells with random cell_info."""
    import torch

    r, e = make_pair(ref, RT, TET, 4, LEG)
    nc = 50_000_000
    ci, ci64 = _cell_info(nc, 11)
    data = torch.randn(nc, 70, generator=torch.Generator(device="cuda").manual_seed(12), device="cuda", dtype=torch.float64)
    idx = torch.from_numpy(_sample(nc, 5000, 4)).cuda()
    before = data[idx].cpu().numpy()
    interior = data[:, 40:].clone()  # the 30 interior dofs never move
    e.T_apply_batch(data, 1, ci)
    assert torch.equal(data[:, 40:], interior)
    del interior
    want = r.T_apply_batch("T_apply", before.copy(), 1, ci64[idx].cpu().numpy().astype(np.uint32))
    got = data[idx].cpu().numpy()
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    e.T_apply_batch(data, 1, ci, op="Tinv_apply")
    assert np.max(np.abs(data[idx].cpu().numpy() - before)) < 1e-11
    chk = torch.randn(nc, 70, generator=torch.Generator(device="cuda").manual_seed(12), device="cuda", dtype=torch.float64)
    assert float((data - chk).abs().max()) < 1e-11
    del data, chk
    _free()


def test_permute_round_trip_50m_cells(ref):
    """BASELINE config 5 (ii): P5 tetrahedron, permute (int32, bit-exact) over 50 M cells."""
    import torch

    r, e = make_pair(ref, P, TET, 5, GLL)
    nc = 50_000_000
    ci, ci64 = _cell_info(nc, 13)
    d = torch.arange(56, dtype=torch.int32, device="cuda").repeat(nc, 1)
    e.permute_batch(d, ci)
    # every row is still a permutation of 0..55: sum and sum of squares per row, then a checksum over all rows
    rows = d.to(torch.int64)
    assert int((rows.sum(dim=1) != 56 * 55 // 2).sum()) == 0
    assert int(((rows * rows).sum(dim=1) != sum(i * i for i in range(56))).sum()) == 0
    del rows
    idx = torch.from_numpy(_sample(nc, 5000, 5)).cuda()
    base = np.tile(np.arange(56, dtype=np.int32), (5000, 1))
    assert np.array_equal(d[idx].cpu().numpy(), r.permute_batch(base, ci64[idx].cpu().numpy().astype(np.uint32)))
    e.permute_batch(d, ci, inverse=True)
    assert bool((d == torch.arange(56, dtype=torch.int32, device="cuda")[None]).all())
    del d
    _free()


def test_tabulate_p1_tri_1m_points(ref):
    """BASELINE config 1: P1 triangle, nd = 1, 1 M points -- the WHOLE table against the compiled reference."""
    from conftest import random_points

    r, e = make_pair(ref, P, 2, 1, GLL)
    x = random_points(2, 1_000_000, seed=42)
    want = r.tabulate(1, x, chunk=65536, nthreads=8)
    got = e.tabulate(1, x)
    assert got.shape == want.shape == (3, 1_000_000, 3, 1)
    for d in range(3):
        assert np.max(np.abs(got[d] - want[d])) <= 1e-12 * np.max(np.abs(want[d]))
    assert np.max(np.abs(got[0].sum(axis=1) - 1)) < 1e-13 and np.max(np.abs(got[1:].sum(axis=2))) < 1e-13


def _jacobians_device(nc, seed):
    import torch

    g = torch.Generator(device="cuda").manual_seed(seed)
    J = torch.eye(3, dtype=torch.float64, device="cuda")[None] + 0.3 * torch.randn(nc, 3, 3, generator=g, device="cuda",
                                                                                 dtype=torch.float64)

    def det_and_inverse(J):
        a, b, c, d, e_, f, g_, h, i = (J[:, r_, c_] for r_ in range(3) for c_ in range(3))
        co = torch.stack([e_ * i - f * h, c * h - b * i, b * f - c * e_, f * g_ - d * i, a * i - c * g_, c * d - a * f,
                          d * h - e_ * g_, b * g_ - a * h, a * e_ - b * d], dim=1)
        det = a * co[:, 0] + b * co[:, 3] + c * co[:, 6]
        return det.contiguous(), (co / det[:, None]).reshape(-1, 3, 3).contiguous()

    detJ, K = det_and_inverse(J)
    J[detJ.abs() < 0.1] = torch.eye(3, dtype=torch.float64, device="cuda")
    detJ, K = det_and_inverse(J)
    return J, detJ, K, g


@pytest.mark.parametrize("family,degree", [(N1E, 3), (RT, 4)])
def test_push_pull_rows45_4m_cells(ref, family, degree):
    """BASELINE config 4 with rows = ndofs of N1curl3 (45, the shape the reference's tests use, test_mappings.py:39-53),
    4 M cells (4.3 GB per array): push_forward AND pull_back sampled against the compiled reference bit for bit, round
    trip over all cells; covariant (N1curl) and contravariant (RT) Piola."""
    import torch

    r, e = make_pair(ref, family, TET, degree, LEG)
    nc, rows = 4_000_000, 45
    J, detJ, K, g = _jacobians_device(nc, 17)
    U = torch.randn(nc, rows, 3, generator=g, device="cuda", dtype=torch.float64)
    u = e.push_forward(U, J, detJ, K)
    back = e.pull_back(u, J, detJ, K)
    assert float(((back - U).abs().amax(dim=(1, 2)) / (1 + U.abs().amax(dim=(1, 2)))).max()) < 1e-9
    idx = torch.from_numpy(_sample(nc, 3000, 6)).cuda()
    a = [v[idx].cpu().numpy() for v in (U, J, detJ, K)]
    ui = u[idx].cpu().numpy()
    assert np.array_equal(ui, r.push_forward(*a))
    assert np.array_equal(back[idx].cpu().numpy(), r.pull_back(ui, *a[1:]))
    # pull_back of independent data (not a round trip)
    v = torch.randn(nc, rows, 3, generator=g, device="cuda", dtype=torch.float64)
    pv = e.pull_back(v, J, detJ, K)
    assert np.array_equal(pv[idx].cpu().numpy(), r.pull_back(v[idx].cpu().numpy(), *a[1:]))
    del J, K, U, u, back, v, pv
    _free()


def test_T_apply_block3_50m_cells(ref):
    """BASELINE config 5 (i) with block size 3 (SURVEY 8d "also n=3"): RT degree 4 tetrahedron, 50 M cells x 210 doubles
    (84 GB in place); sample against the compiled reference, interior dofs untouched, Tinv_apply restores the input."""
    import torch

    r, e = make_pair(ref, RT, TET, 4, LEG)
    nc, bs, chunk = 50_000_000, 3, 5_000_000
    ci, ci64 = _cell_info(nc, 21)
    data = torch.empty(nc, 70 * bs, device="cuda", dtype=torch.float64)

    def regenerate(c):
        g = torch.Generator(device="cuda").manual_seed(1000 + c)
        return torch.randn(chunk, 70 * bs, generator=g, device="cuda", dtype=torch.float64)

    for c in range(nc // chunk):
        data[c * chunk:(c + 1) * chunk] = regenerate(c)
    idx = torch.from_numpy(_sample(nc, 4000, 8)).cuda()
    before = data[idx].cpu().numpy()
    e.T_apply_batch(data, bs, ci)
    want = r.T_apply_batch("T_apply", before.copy(), bs, ci64[idx].cpu().numpy().astype(np.uint32))
    got = data[idx].cpu().numpy()
    assert np.max(np.abs(got - want)) <= 1e-12 * np.max(np.abs(want))
    for c in range(nc // chunk):  # the 30 interior dofs (x 3 block entries) never move
        assert torch.equal(data[c * chunk:(c + 1) * chunk, 40 * bs:], regenerate(c)[:, 40 * bs:])
    e.T_apply_batch(data, bs, ci, op="Tinv_apply")
    for c in range(nc // chunk):
        assert float((data[c * chunk:(c + 1) * chunk] - regenerate(c)).abs().max()) < 1e-11
    del data
    _free()
