"""End-to-end use of the path with a natively built element and no oracle in the loop: the basis functions of an
N1curl / RT element, tabulated at the element's own interpolation points, pushed to random physical cells with the
element's Piola map (one slab for all cells), pulled back, interpolated, must come back as unit coefficient vectors;
DOF transformations applied and undone in between leave them unchanged."""

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("family,degree", [(3, 3), (2, 2), (3, 1)])  # N1curl 3, RT 2, N1curl 1 on the tetrahedron
def test_tabulate_push_pull_interpolate_round_trip(family, degree):
    import torch

    import basix_b200 as bx

    e = bx.create_element(family, 3, degree, 11)
    pts = e.points
    npts, n = pts.shape[0], e.dim
    phi = e.tabulate(0, pts)[0]  # (npts, ndofs, 3)
    rng = np.random.default_rng(degree)
    nc = 4099
    J = np.eye(3)[None] + 0.25 * rng.standard_normal((nc, 3, 3))
    detJ, K = np.linalg.det(J), np.linalg.inv(J)
    dev = [torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (J, detJ, K)]
    ci = torch.from_numpy(rng.integers(0, 2 ** 30, nc)).cuda().to(torch.uint32)
    for dof in (0, n // 2, n - 1):
        U0 = torch.from_numpy(np.ascontiguousarray(phi[:, dof, :])).cuda()  # the dof's basis function at the points
        u = e.push_forward_broadcast(U0, *dev)  # (nc, npts, 3) physical values on every cell
        back = e.pull_back(u, *dev)
        coef = e.interpolate_batch(back.reshape(nc, npts * 3), layout="point-major")  # (nc, ndofs)
        want = torch.zeros(nc, n, dtype=torch.float64, device="cuda")
        want[:, dof] = 1.0
        assert float((coef - want).abs().max()) < 1e-9
        # transform to the physical dof orientation and back
        w = coef.clone()
        e.T_apply_batch(w, 1, ci, op="Tt_inv_apply")
        e.T_apply_batch(w, 1, ci, op="Tt_apply")
        assert float((w - coef).abs().max()) < 1e-12
