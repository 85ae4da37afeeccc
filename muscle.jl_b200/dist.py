"""Multi-GPU `binary_einsum`: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on
the box, gloo in CPU tests) for the plumbing.

Mirrors the reference's only distributed path, Dagger block sharding
(ext/MuscleDaggerExt/binary_einsum.jl:64-119):
  * splitting a FREE or BATCH index gives independent output blocks — Dagger's loop over output blocks
    (:88-105). Here: each rank contracts its slab, no collective on the data path.
  * splitting a SUMMED index gives full-size partial outputs that are add-reduced — Dagger's
    `treereduce(AddComputeOp, …)` over the summed blocks (:107-115). Here: `all_reduce(SUM)`.
The per-shard contraction is the ordinary single-GPU `binary_einsum` (Dagger re-enters
`Muscle.binary_einsum` per chunk too, :60-62).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .backend import BackendB200
from .einsum import binary_einsum, flatten_labels
from .tensor import B200Array, Tensor, _as_index_list


def plan_shard(a: Tensor, b: Tensor, inds_c, nranks: int, rank: int, prefer_sum: bool = False):
    """Which index this rank's work is cut along. Returns (kind, index, begin, end, needs_allreduce);
    kind is one of _lib.SHARD_NONE / SHARD_FREE / SHARD_BATCH / SHARD_SUM."""
    inds_c = _as_index_list(inds_c)
    ma, mb, mc = flatten_labels(a.inds, b.inds, inds_c)
    info = _lib.shard_plan(mc, ma, a.shape, mb, b.shape, nranks, rank, prefer_sum)
    index = None
    if info.kind != _lib.SHARD_NONE:
        for lab, m in list(zip(a.inds, ma)) + list(zip(b.inds, mb)):
            if m == info.mode:
                index = lab
                break
    return info.kind, index, int(info.begin), int(info.end), bool(info.needs_allreduce)


def local_slab(t: Tensor, index, begin: int, end: int) -> Tensor:
    """This rank's slab of a HOST tensor along `index` (the data-distribution step; tensors that do not
    carry the index are replicated). Device tensors are expected to be created per rank already."""
    if index not in t.inds:
        return t
    if t.on_device:
        raise _lib.ArgumentError("local_slab works on host tensors; create device slabs per rank")
    sl = [slice(None)] * t.ndim
    sl[t.dim(index)] = slice(begin, end)
    return Tensor(_lib.fortran(t.data[tuple(sl)]), t.inds)


def _default_contract(inds_c, a, b):
    return binary_einsum(BackendB200(), inds_c, a, b)


def sharded_binary_einsum(a: Tensor, b: Tensor, inds_c, *, group=None, prefer_sum=False, gather=False,
                          contract=_default_contract):
    """Contract replicated host (or per-rank device) operands across the ranks of `group`.

    Returns (c_local, info). For a free/batch shard `c_local` is this rank's slab of C (or, with
    gather=True, the full C assembled by all_gather along the split index); for a summed-index slice
    it is the full C after all_reduce(SUM). `contract` exists so the CPU (gloo) tests can exercise this
    host logic without a GPU; the product default is BackendB200.
    """
    import torch
    import torch.distributed as dist

    inds_c = _as_index_list(inds_c)
    nranks = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    kind, index, begin, end, needs_allreduce = plan_shard(a, b, inds_c, nranks, rank, prefer_sum)
    if kind == _lib.SHARD_NONE:
        return contract(inds_c, a, b), (kind, index, begin, end)     # replicas only
    a_loc = local_slab(a, index, begin, end)
    b_loc = local_slab(b, index, begin, end)
    c_loc = contract(inds_c, a_loc, b_loc)
    if needs_allreduce:
        c_loc = all_reduce_sum(c_loc, group)
    elif gather:
        c_loc = all_gather_along(c_loc, index, group)
    return c_loc, (kind, index, begin, end)


def _as_torch_view(c: Tensor):
    """A torch tensor aliasing C's storage (device: the owning uint8 buffer reinterpreted; host: numpy)."""
    import torch
    if c.on_device:
        owner = c.data._owner
        if owner is None:
            raise _lib.B200Error("device array is not backed by a torch allocation; cannot hand it to NCCL")
        real = torch.float32 if c.dtype in (np.dtype(np.float32), np.dtype(np.complex64)) else torch.float64
        return owner[: c.data.nbytes].view(real)
    arr = c.data
    flat = arr.reshape(-1, order="F") if arr.ndim else arr.reshape(1)
    if not (arr.flags.f_contiguous or arr.ndim <= 1):
        raise _lib.ArgumentError("host C must be column-major contiguous")
    real = np.float32 if arr.dtype in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64
    return torch.from_numpy(flat.view(real))


def all_reduce_sum(c: Tensor, group=None) -> Tensor:
    """Σ over ranks of the partial outputs, in place (a sum does not care about re/im interleaving, so the
    buffer is reduced as plain reals)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if c.on_device:
            _lib.Handle.get(c.data.device)       # make sure our launches and NCCL share torch's current stream
        dist.all_reduce(_as_torch_view(c), op=dist.ReduceOp.SUM, group=group)
    return c


def all_gather_along(c: Tensor, index, group=None) -> Tensor:
    """Assemble the full C from per-rank slabs along `index` (host tensors; used by tests and by callers
    that want C replicated — the timed sharded path keeps C sharded)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return c
    host = c.to_host()
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, host.data, group=group)
    full = np.concatenate(parts, axis=host.dim(index))
    return Tensor(_lib.fortran(full), host.inds)
