"""World-size-2 `gloo` test (CPU) of the multi-GPU host logic: slicing rule, slab-wise gather of a point-sharded
tabulate result, gather of cell-major results, in-place sharded DOF transformation.  The per-slice compute is
done by the oracle restatement standing in for the CUDA element (same method names), so no GPU is needed; on a
GPU box the same functions run with the real element and NCCL (bench.py --gpus N uses the same slicing)."""

import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_bounds_cover_and_are_contiguous():
    from basix_b200.sharding import shard_bounds, shard_sizes

    for n in (0, 1, 7, 8, 1000003):
        for world in (1, 2, 3, 8):
            edges = [shard_bounds(n, world, r) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[r][1] == edges[r + 1][0] for r in range(world - 1))
            assert sum(shard_sizes(n, world)) == n and max(shard_sizes(n, world)) - min(shard_sizes(n, world)) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


class _OracleElement:
    """Stand-in with the product element's method names, computing through oracle/restate.py (torch CPU in/out)."""

    def __init__(self, name):
        from oracle import restate as R

        arrs = np.load(os.path.join(ROOT, "tests", "golden", "elements", name + ".npz"))
        self.e = R.RestatedElement(arrs)
        self.M = arrs["interpolation_matrix"]

    def tabulate(self, nd, x):
        import torch

        return torch.from_numpy(self.e.tabulate(nd, x.numpy()))

    def push_forward(self, U, J, detJ, K):
        import torch

        return torch.from_numpy(self.e.push_forward(U.numpy(), J.numpy(), detJ.numpy(), K.numpy()))

    def pull_back(self, u, J, detJ, K):
        import torch

        return torch.from_numpy(self.e.pull_back(u.numpy(), J.numpy(), detJ.numpy(), K.numpy()))

    def T_apply_batch(self, data, bs, ci, op="T_apply"):
        self.e.T_apply_batch(op, data.numpy(), bs, ci.numpy())

    def push_forward_broadcast(self, U, J, detJ, K):
        import torch

        rep = np.ascontiguousarray(np.broadcast_to(U.numpy(), (J.shape[0],) + tuple(U.shape)))
        return torch.from_numpy(self.e.push_forward(rep, J.numpy(), detJ.numpy(), K.numpy()))

    def interpolate_batch(self, values, layout="component-major"):
        import torch

        assert layout == "component-major"
        return torch.from_numpy(values.numpy() @ self.M.T)

    def interpolate_pulled_batch(self, u, J, detJ, K):
        U = self.pull_back(u, J, detJ, K).numpy()  # (nc, npts, vs) -> the reference's component-major columns
        return self.interpolate_batch(__import__("torch").from_numpy(np.ascontiguousarray(U.transpose(0, 2, 1)).reshape(len(U), -1)))

    def physical_basis(self, X, coords=None, J=None, detJ=None, K=None, cell_info=None, op="T_apply"):
        import torch

        from oracle import restate as R

        U = self.e.tabulate(0, X.numpy())[0]  # (npts, ndofs, vs)
        npts, nd, vs = U.shape
        Jn, dn, Kn = R.geometry(coords.numpy())
        nc = len(Jn)
        rep = np.ascontiguousarray(np.broadcast_to(U.reshape(1, npts, nd * vs), (nc, npts, nd * vs))).reshape(nc * npts, nd * vs)
        self.e.T_apply_batch(op, rep, vs, np.repeat(cell_info.numpy(), npts))
        return torch.from_numpy(self.e.push_forward(rep.reshape(nc, npts * nd, vs), Jn, dn, Kn).reshape(nc, npts, nd, -1))


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch
    import torch.distributed as dist

    from basix_b200 import sharding

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(0)  # same inputs on every rank
        # tabulate: 101 points (odd -> ragged slices), gathered to all and to rank 1 only
        e = _OracleElement("P3_triangle_gll")
        x = rng.random((101, 2)) / 2
        full = e.tabulate(1, torch.from_numpy(x))
        got = sharding.tabulate_sharded(e, 1, torch.from_numpy(x), gather=True)
        assert torch.equal(got, full)
        only1 = sharding.tabulate_sharded(e, 1, torch.from_numpy(x), gather=True, dst=1)
        assert (only1 is None) == (rank != 1) and (rank != 1 or torch.equal(only1, full))
        local = sharding.tabulate_sharded(e, 1, torch.from_numpy(x))
        lo, hi = sharding.shard_bounds(101, world, rank)
        assert torch.equal(local, full[:, lo:hi])
        # maps: 33 cells
        e = _OracleElement("N1curl3_tetrahedron_legendre")
        J = np.eye(3)[None] + 0.3 * rng.standard_normal((33, 3, 3))
        args = [torch.from_numpy(np.ascontiguousarray(a)) for a in
                (rng.standard_normal((33, 4, 3)), J, np.linalg.det(J), np.linalg.inv(J))]
        assert torch.equal(sharding.map_sharded(e, False, *args, gather=True), e.push_forward(*args))
        assert torch.equal(sharding.map_sharded(e, True, *args, gather=True), e.pull_back(*args))
        # one slab of reference values for all cells; batched interpolation of 33 cells
        U0 = torch.from_numpy(rng.standard_normal((4, 3)))
        want = e.push_forward_broadcast(U0, *args[1:])
        assert torch.equal(sharding.push_forward_broadcast_sharded(e, U0, *args[1:], gather=True), want)
        vals = torch.from_numpy(rng.standard_normal((33, e.M.shape[1])))
        # (the stand-in's dgemm blocks a slice differently from the whole array: compare to rounding)
        assert torch.allclose(sharding.interpolate_sharded(e, vals, gather=True), e.interpolate_batch(vals), rtol=0, atol=1e-12)
        # the fused calls: pull_back + interpolation, and tabulate-once -> T_apply -> push_forward from vertex coordinates
        npts = e.M.shape[1] // 3
        u = torch.from_numpy(rng.standard_normal((33, npts, 3)))
        assert torch.allclose(sharding.interpolate_pulled_sharded(e, u, *args[1:], gather=True),
                              e.interpolate_pulled_batch(u, *args[1:]), rtol=0, atol=1e-11)
        verts = np.zeros((4, 3))
        verts[1:] = np.eye(3)
        coords = torch.from_numpy(verts[None] + 0.1 * rng.standard_normal((33, 4, 3)))
        ci33 = torch.from_numpy(rng.integers(0, 2 ** 18, 33).astype(np.int64))
        X = torch.from_numpy(rng.random((3, 3)) / 4)
        want = e.physical_basis(X, coords=coords, cell_info=ci33)
        assert torch.equal(sharding.physical_basis_sharded(e, X, coords=coords, cell_info=ci33, gather=True), want)
        lo, hi = sharding.shard_bounds(33, world, rank)
        assert torch.equal(sharding.physical_basis_sharded(e, X, coords=coords, cell_info=ci33), want[lo:hi])
        # in-place sharded T_apply, then exchange of the slices
        e = _OracleElement("RT2_tetrahedron_legendre")
        data = torch.from_numpy(rng.standard_normal((21, 15)))
        ci = torch.from_numpy(rng.integers(0, 2 ** 30, 21).astype(np.int64))
        want = data.clone()
        e.T_apply_batch(want, 1, ci)
        lo, hi = sharding.T_apply_sharded(e, data, 1, ci)
        assert torch.equal(data[lo:hi], want[lo:hi])
        assert torch.equal(sharding.gather_cells(data[lo:hi], 21), want)
        # equal slices take the copy-free path (all_gather_into_tensor / gather straight into the result)
        e = _OracleElement("P3_triangle_gll")
        x = rng.random((100, 2)) / 2
        full = e.tabulate(1, torch.from_numpy(x))
        assert torch.equal(sharding.tabulate_sharded(e, 1, torch.from_numpy(x), gather=True), full)
        only0 = sharding.tabulate_sharded(e, 1, torch.from_numpy(x), gather=True, dst=0)
        assert (only0 is None) == (rank != 0) and (rank != 0 or torch.equal(only0, full))
        pre = torch.empty(20, 15, dtype=torch.float64)
        blk = torch.full((10, 15), float(rank), dtype=torch.float64)
        assert sharding.gather_cells(blk, 20, out=pre) is pre and torch.equal(pre[::10, 0], torch.tensor([0.0, 1.0]))
        # a sub-group: slice bounds come from the rank and size IN THE GROUP (here one rank per group: the whole array)
        groups = [dist.new_group([r]) for r in range(world)]
        e = _OracleElement("RT2_tetrahedron_legendre")
        data = torch.from_numpy(rng.standard_normal((21, 15)))
        ci = torch.from_numpy(rng.integers(0, 2 ** 30, 21).astype(np.int64))
        want = data.clone()
        e.T_apply_batch(want, 1, ci)
        assert sharding.T_apply_sharded(e, data, 1, ci, group=groups[rank]) == (0, 21)
        assert torch.equal(data, want)
        with open(os.path.join(out_dir, f"ok{rank}"), "w") as f:
            f.write("ok")
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo(tmp_path):
    import torch.multiprocessing as mp

    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    assert sorted(os.listdir(tmp_path)) == ["ok0", "ok1"]
