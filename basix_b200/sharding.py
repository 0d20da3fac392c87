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


def gather_cells(local, n_total: int, dst: int | None = None, group=None):
    """Gather a unit-major array (first axis = this rank's slice of the ``n_total`` units) onto every rank
    (``dst=None``) or onto rank ``dst`` only (other ranks get ``None``).  Slices may differ in length by one."""
    import torch

    dist, world, rank = _dist(group)
    sizes = shard_sizes(n_total, world)
    if local.shape[0] != sizes[rank]:
        raise ValueError(f"rank {rank}: expected {sizes[rank]} units, got {local.shape[0]}")
    local = local.contiguous()
    if world == 1:
        return local
    # collectives need equal shapes: pad the (at most one unit) shorter slices, trim after the exchange
    smax = max(sizes)
    if local.shape[0] < smax:
        pad = torch.zeros((smax - local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype, device=local.device)
        local = torch.cat([local, pad], dim=0)
    parts = [torch.empty_like(local) for _ in sizes]
    if dst is None:
        dist.all_gather(parts, local, group=group)
    else:
        dist.gather(local, parts if rank == dst else None, dst=dst, group=group)
        if rank != dst:
            return None
    return torch.cat([p[:s] for p, s in zip(parts, sizes)], dim=0)


def gather_tabulate(local_basis, npts_total: int, dst: int | None = None, group=None):
    """Gather a point-sharded tabulate result ``[nderivs][npts_local][ndofs][vs]`` into
    ``[nderivs][npts_total][ndofs][vs]``: one gather per derivative slab (each slab is contiguous per rank)."""
    import torch

    slabs = [gather_cells(local_basis[d], npts_total, dst, group) for d in range(local_basis.shape[0])]
    if any(s is None for s in slabs):
        return None
    return torch.stack(slabs, dim=0)


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


def T_apply_sharded(element, data, block_size: int, cell_info, op: str = "T_apply"):
    """In-place DOF transformation of this rank's slice of the cells of ``data[ncells][ndofs*block_size]``;
    returns the slice bounds.  (In place: there is nothing to gather unless the caller wants the slices
    exchanged, for which :func:`gather_cells` applies to ``data[lo:hi]``.)"""
    _, world, rank = _dist(group=None)
    lo, hi = shard_bounds(data.shape[0], world, rank)
    element.T_apply_batch(data[lo:hi], block_size, cell_info[lo:hi], op=op)
    return lo, hi
