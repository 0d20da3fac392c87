"""Multi-GPU plumbing for the hot path: one process per GPU, independent contiguous slices, no data-path
collective.

Points (``tabulate``) and cells (``push_forward`` / ``pull_back`` / ``T_apply`` / ``permute``) are independent
units; element data is replicated on every rank.  Rank ``r`` of ``G`` works on units ``[r*N//G, (r+1)*N//G)``.
A collective is used only when a caller wants the full result on one device or on all of them:

* cell-major results (maps, DOF transformations) are one contiguous block per rank -> a single gather;
* the tabulate result is ``[deriv][point][dof][vs]``, so a point-sharded result is ``nderivs`` separate
  contiguous chunks per rank -> gathered slab by slab.

The collectives are ``torch.distributed`` calls (NCCL over NVLink when the tensors are CUDA tensors, gloo in
the CPU tests); no arithmetic happens here.
"""

from __future__ import annotations


def shard_bounds(n: int, world_size: int, rank: int) -> tuple[int, int]:
    """Contiguous slice of ``n`` units owned by ``rank``: ``[rank*n//G, (rank+1)*n//G)``."""
    if world_size < 1 or not 0 <= rank < world_size:
        raise ValueError(f"bad rank {rank} for world size {world_size}")
    return rank * n // world_size, (rank + 1) * n // world_size


def shard_sizes(n: int, world_size: int) -> list[int]:
    return [shard_bounds(n, world_size, r)[1] - shard_bounds(n, world_size, r)[0] for r in range(world_size)]


def _dist(group=None):
    import torch.distributed as dist

    if not dist.is_available() or not dist.is_initialized():
        raise RuntimeError("basix_b200.sharding: torch.distributed is not initialised")
    return dist, dist.get_world_size(group), dist.get_rank(group)


def _gather_into(out, local, sizes, rank, dst, group):
    """Collect the per-rank blocks ``local`` into ``out`` (contiguous, first axis = all units in rank order) with no
    intermediate copy when the slices are equal: ``all_gather_into_tensor`` writes every rank's block straight into
    its place in ``out`` (NCCL all-gather over NVLink on CUDA tensors); with ``dst`` the root receives each block
    straight into its view of ``out``.  Ragged slices (lengths differing by one unit) go through padded blocks."""
    import torch

    dist, world, _ = _dist(group)
    equal = min(sizes) == max(sizes)
    if dst is None:
        if equal:
            dist.all_gather_into_tensor(out, local, group=group)
            return
        smax = max(sizes)
        padded = torch.zeros((smax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        padded[: local.shape[0]] = local
        flat = torch.empty((world * smax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        dist.all_gather_into_tensor(flat, padded, group=group)
        lo = 0
        for r, s in enumerate(sizes):
            out[lo:lo + s] = flat[r * smax:r * smax + s]
            lo += s
        return
    dst_global = dist.get_global_rank(group, dst) if group is not None else dst
    if equal:
        views = None
        if rank == dst:
            views, lo = [], 0
            for s in sizes:
                views.append(out[lo:lo + s])
                lo += s
        dist.gather(local, views, dst=dst_global, group=group)
        return
    smax = max(sizes)
    padded = torch.zeros((smax,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    padded[: local.shape[0]] = local
    parts = [torch.empty_like(padded) for _ in sizes] if rank == dst else None
    dist.gather(padded, parts, dst=dst_global, group=group)
    if rank == dst:
        lo = 0
        for p, s in zip(parts, sizes):
            out[lo:lo + s] = p[:s]
            lo += s


def gather_cells(local, n_total: int, dst: int | None = None, group=None, out=None):
    """Gather a unit-major array (first axis = this rank's slice of the ``n_total`` units) onto every rank
    (``dst=None``) or onto rank ``dst`` of the group only (other ranks get ``None``).  Slices may differ in length by
    one.  ``out`` (optional, contiguous ``[n_total, ...]``) receives the result in place."""
    import torch

    dist, world, rank = _dist(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: expected {sizes[rank]} units, got {local.shape[0]}")
    local = local.contiguous()
    if world == 1:
        if out is not None:
            out.copy_(local)
            return out
        return local
    need = dst is None or rank == dst
    if need and out is None:
        out = torch.empty((n_total,) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
    if need and (not out.is_contiguous() or tuple(out.shape) != (n_total,) + tuple(local.shape[1:])):
        raise ValueError("gather_cells: out must be contiguous with shape (n_total, ...)")
    _gather_into(out, local, sizes, rank, dst, group)
    return out if need else None


def gather_tabulate(local_basis, npts_total: int, dst: int | None = None, group=None):
    """Gather a point-sharded tabulate result ``[nderivs][npts_local][ndofs][vs]`` into
    ``[nderivs][npts_total][ndofs][vs]``: the result is allocated once and every derivative slab -- one contiguous
    run per rank on both sides -- is gathered straight into its place (no concatenation or stacking copies)."""
    import torch

    dist, world, rank = _dist(group)
    nder = local_basis.shape[0]
    need = dst is None or rank == dst
    out = (torch.empty((nder, npts_total) + tuple(local_basis.shape[2:]), dtype=local_basis.dtype,
                       device=local_basis.device) if need else None)
    for d in range(nder):
        gather_cells(local_basis[d], npts_total, dst, group, out=out[d] if need else None)
    return out


def tabulate_sharded(element, nd: int, x, gather: bool = False, dst: int | None = None, group=None):
    """``element.tabulate(nd, x[lo:hi])`` on this rank's slice of the points ``x`` (the same full array on every
    rank); with ``gather=True`` the slabs are collected as in :func:`gather_tabulate`."""
    _, world, rank = _dist(group)
    lo, hi = shard_bounds(x.shape[0], world, rank)
    local = element.tabulate(nd, x[lo:hi])
    return gather_tabulate(local, x.shape[0], dst, group) if gather else local


def map_sharded(element, pull: bool, U, J, detJ, K, gather: bool = False, dst: int | None = None, group=None):
    """push_forward (``pull=False``) / pull_back on this rank's slice of the cells."""
    _, world, rank = _dist(group)
    lo, hi = shard_bounds(U.shape[0], world, rank)
    fn = element.pull_back if pull else element.push_forward
    local = fn(U[lo:hi], J[lo:hi], detJ[lo:hi], K[lo:hi])
    return gather_cells(local, U.shape[0], dst, group) if gather else local


def push_forward_broadcast_sharded(element, U, J, detJ, K, gather: bool = False, dst: int | None = None, group=None):
    """``push_forward_broadcast`` (one slab of reference values ``U`` for all cells) on this rank's slice of the cells."""
    _, world, rank = _dist(group)
    lo, hi = shard_bounds(J.shape[0], world, rank)
    local = element.push_forward_broadcast(U, J[lo:hi], detJ[lo:hi], K[lo:hi])
    return gather_cells(local, J.shape[0], dst, group) if gather else local


def interpolate_sharded(element, values, layout: str = "component-major", gather: bool = False, dst: int | None = None,
                        group=None):
    """``interpolate_batch`` on this rank's slice of the cells of ``values[ncells][value_size * npts]``."""
    _, world, rank = _dist(group)
    lo, hi = shard_bounds(values.shape[0], world, rank)
    local = element.interpolate_batch(values[lo:hi], layout=layout)
    return gather_cells(local, values.shape[0], dst, group) if gather else local


def interpolate_pulled_sharded(element, u, J, detJ, K, gather: bool = False, dst: int | None = None, group=None):
    """``interpolate_pulled_batch`` (pull_back fused into the interpolation) on this rank's slice of the cells."""
    _, world, rank = _dist(group)
    lo, hi = shard_bounds(u.shape[0], world, rank)
    local = element.interpolate_pulled_batch(u[lo:hi], J[lo:hi], detJ[lo:hi], K[lo:hi])
    return gather_cells(local, u.shape[0], dst, group) if gather else local


def physical_basis_sharded(element, X, coords=None, J=None, detJ=None, K=None, cell_info=None, op: str = "T_apply",
                           gather: bool = False, dst: int | None = None, group=None):
    """``physical_basis`` (tabulate once, T_apply and push_forward fused) on this rank's slice of the cells; the points
    ``X`` are the same on every rank, every per-cell array is sliced."""
    _, world, rank = _dist(group)
    first = next(a for a in (coords, J, K, detJ, cell_info) if a is not None)
    nc = first.shape[0]
    lo, hi = shard_bounds(nc, world, rank)
    cut = lambda a: None if a is None else a[lo:hi]  # noqa: E731
    local = element.physical_basis(X, coords=cut(coords), J=cut(J), detJ=cut(detJ), K=cut(K), cell_info=cut(cell_info), op=op)
    return gather_cells(local, nc, dst, group) if gather else local


def T_apply_sharded(element, data, block_size: int, cell_info, op: str = "T_apply", group=None):
    """In-place DOF transformation of this rank's slice of the cells of ``data[ncells][ndofs*block_size]``;
    returns the slice bounds.  (In place: there is nothing to gather unless the caller wants the slices
    exchanged, for which :func:`gather_cells` applies to ``data[lo:hi]``.)"""
    _, world, rank = _dist(group)
    lo, hi = shard_bounds(data.shape[0], world, rank)
    element.T_apply_batch(data[lo:hi], block_size, cell_info[lo:hi], op=op)
    return lo, hi
