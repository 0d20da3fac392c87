"""Device-pointer calls are stream-ordered and launch-only: each one can be captured into a CUDA graph and replayed on
new data (bx_element_prepare / element creation do the one-off uploads beforehand)."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _capture(fn):
    """Run fn() eagerly once (warm-up), capture it, return (graph, captured result)."""
    import torch

    fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        out = fn()
    return g, out


def _simplex_points(n, tdim, seed):
    import torch

    gen = torch.Generator(device="cuda")
    gen.manual_seed(seed)
    return torch.rand(n, tdim, device="cuda", dtype=torch.float64, generator=gen) / (tdim + 0.5)


@pytest.mark.parametrize("family,cell,degree,variant,nd", [(1, 3, 5, 2, 2), (1, 2, 1, 2, 1), (1, 5, 4, 2, 1), (3, 3, 3, 11, 1)])
def test_tabulate_in_a_cuda_graph(family, cell, degree, variant, nd):
    import torch

    import basix_b200 as bx

    e = bx.create_element(family, cell, degree, variant)
    tdim = {2: 2, 3: 3, 5: 3}[cell]
    x = _simplex_points(5000 + 3, tdim, 1)
    g, out = _capture(lambda: e.tabulate(nd, x))
    x.copy_(_simplex_points(5000 + 3, tdim, 2))
    g.replay()
    torch.cuda.synchronize()
    assert torch.equal(out, e.tabulate(nd, x))


def test_maps_transformations_and_fused_basis_in_a_cuda_graph():
    import torch

    import basix_b200 as bx

    n1 = bx.create_element(3, 3, 3, 11)
    rt = bx.create_element(2, 3, 4, 11)
    p5 = bx.create_element(1, 3, 5, 2)
    nc = 3000 + 7
    gen = torch.Generator(device="cuda")
    gen.manual_seed(3)

    def rnd(*shape):
        return torch.randn(*shape, device="cuda", dtype=torch.float64, generator=gen)

    J = torch.eye(3, device="cuda", dtype=torch.float64)[None] + 0.2 * rnd(nc, 3, 3)
    detJ, K = torch.linalg.det(J), torch.linalg.inv(J).contiguous()
    U = rnd(nc, 45, 3)
    verts = torch.tensor([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1.0]], device="cuda", dtype=torch.float64)
    coords = (verts[None] + 0.1 * rnd(nc, 4, 3)).contiguous()
    ci = torch.randint(0, 2 ** 18, (nc,), device="cuda", dtype=torch.int64, generator=gen).to(torch.int32)
    X = _simplex_points(2, 3, 4)
    data = rnd(nc, 70)
    perm = torch.arange(56, device="cuda", dtype=torch.int32).repeat(nc, 1).contiguous()
    vals = rnd(nc, n1.interpolation_matrix.shape[1])

    def step():
        u = n1.push_forward(U, J, detJ, K)
        back = n1.pull_back(u, J, detJ, K)
        d = data.clone()
        rt.T_apply_batch(d, 1, ci)
        q = perm.clone()
        p5.permute_batch(q, ci)
        phys = rt.physical_basis(X, coords=coords, cell_info=ci)
        geo = bx.geometry(coords)
        coef = n1.interpolate_batch(vals)
        return u, back, d, q, phys, geo[0], geo[1], geo[2], coef

    g, outs = _capture(step)
    # new inputs in the same buffers
    U.copy_(rnd(nc, 45, 3))
    data.copy_(rnd(nc, 70))
    coords.copy_((verts[None] + 0.1 * rnd(nc, 4, 3)))
    ci.copy_(torch.randint(0, 2 ** 18, (nc,), device="cuda", dtype=torch.int64, generator=gen).to(torch.int32))
    vals.copy_(rnd(*vals.shape))
    g.replay()
    torch.cuda.synchronize()
    want = step()
    for a, b in zip(outs, want):
        assert torch.equal(a, b)
