"""Blocked tensors — the Dagger bridge.

The reference's only distributed path wraps `Dagger.DArray`s in `Tensor`s and contracts them block by block
(ext/MuscleDaggerExt/binary_einsum.jl): `Dagger.stage(::BinaryEinsum)` (:64-119) builds one output chunk per block of
the output grid, each the add-tree reduction (`treereduce(AddComputeOp)`, :107-115) over the summed-index blocks of
`task_binary_einsum` (:60-62) — an ordinary `binary_einsum` of two chunks that re-enters Muscle's dispatch.

Julia and Dagger cannot run here, so this module is the host-side mirror of that stage: `BlockArray` stands in for a
`DArray` with `Dagger.Blocks` partitioning, `distribute` for `Dagger.distribute`, `BackendBlocks` for `BackendDagger`.
Every chunk contraction goes back through `binary_einsum` — device chunks (`B200Array`) therefore run on BackendB200,
which is exactly what the one-line Domain rule in `julia/MuscleB200.jl` gives Dagger's `task_binary_einsum`.
The reduction over summed blocks stays on the device: every pair is contracted straight into its slot of one staging
buffer (`binary_einsum!` into a view) and `mb200_reduce_slots` adds the slots in order (one pass, no add-tree of
temporaries). Host chunks (numpy) are summed on the host in Dagger's tree order; they need a backend override
(`with_backend(f, BackendB200())`), as host arrays always do in this package.

Multi-GPU: a `BlockArray` lives on one device; several GPUs are driven by the multi-process sharder (`dist.py`).
"""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import _lib
from ._lib import ArgumentError
from .tensor import B200Array, Index, Tensor


class Blocks:
    """`Dagger.Blocks(bs...)`: one block extent per dimension."""

    def __init__(self, *blocksize):
        if len(blocksize) == 1 and isinstance(blocksize[0], (tuple, list)):
            blocksize = tuple(blocksize[0])
        if any(int(b) <= 0 for b in blocksize):
            raise ArgumentError("block sizes must be positive")
        self.blocksize = tuple(int(b) for b in blocksize)

    def __repr__(self):
        return f"Blocks{self.blocksize}"


class BlockArray:
    """Stand-in for a `Dagger.DArray` partitioned by `Dagger.Blocks`: an N-d grid of equally sized chunks."""

    _is_block_array = True

    def __init__(self, chunks: np.ndarray, blocksize, dtype):
        self.chunks = chunks                       # object array, one chunk (numpy or B200Array) per block
        self.blocksize = tuple(int(b) for b in blocksize)
        self.dtype = np.dtype(dtype)
        if chunks.ndim != len(self.blocksize):
            raise ArgumentError("one block size per dimension")
        self.shape = tuple(g * b for g, b in zip(chunks.shape, self.blocksize))

    @property
    def ndim(self) -> int:
        return len(self.shape)

    @property
    def partitioning(self) -> Blocks:
        return Blocks(*self.blocksize)

    @property
    def on_device(self) -> bool:
        return self.chunks.size > 0 and isinstance(self.chunks.flat[0], B200Array)

    def domainchunks(self):
        """Shapes of the chunks, in grid order (`Dagger.domainchunks`)."""
        return [tuple(c.shape) for c in self.chunks.flat]

    def collect(self) -> np.ndarray:
        """`collect(::DArray)`: the dense host array."""
        out = np.empty(self.shape, dtype=self.dtype, order="F")
        for idx in np.ndindex(*self.chunks.shape):
            c = self.chunks[idx]
            sl = tuple(slice(i * b, (i + 1) * b) for i, b in zip(idx, self.blocksize))
            out[sl] = c.to_host() if isinstance(c, B200Array) else c
        return out

    def __repr__(self):
        return f"BlockArray(shape={self.shape}, blocks={self.blocksize}, dtype={self.dtype}, device={self.on_device})"


def distribute(data, blocks: Blocks, device=None) -> BlockArray:
    """`Dagger.distribute(data, Blocks(...))`. `device=None` keeps host chunks; an int uploads every chunk to that GPU.
    The extents must be multiples of the block sizes (Dagger.stage divides with `÷`, binary_einsum.jl:68-69)."""
    data = np.asarray(data)
    bs = blocks.blocksize
    if len(bs) != data.ndim:
        raise ArgumentError(f"{len(bs)} block sizes for a {data.ndim}-d array")
    if any(e % b for e, b in zip(data.shape, bs)):
        raise ArgumentError(f"extents {data.shape} are not multiples of the block sizes {bs}")
    grid = tuple(e // b for e, b in zip(data.shape, bs))
    chunks = np.empty(grid, dtype=object)
    for idx in np.ndindex(*grid):
        sl = tuple(slice(i * b, (i + 1) * b) for i, b in zip(idx, bs))
        c = _lib.fortran(data[sl])
        chunks[idx] = B200Array.from_host(c, device) if device is not None else c
    return BlockArray(chunks, bs, data.dtype)


def _to_block_array(x) -> BlockArray:
    """`Dagger._to_darray` (binary_einsum.jl:31-33): a plain array becomes a single-block BlockArray."""
    if isinstance(x, BlockArray):
        return x
    chunks = np.empty((1,) * x.ndim, dtype=object)
    chunks[(0,) * x.ndim] = x
    return BlockArray(chunks, tuple(x.shape), x.dtype)


def _tree_sum(parts):
    """`Dagger.treereduce(AddComputeOp, parts)`: pairwise add tree (host chunks)."""
    parts = list(parts)
    while len(parts) > 1:
        nxt = [parts[i] + parts[i + 1] for i in range(0, len(parts) - 1, 2)]
        if len(parts) % 2:
            nxt.append(parts[-1])
        parts = nxt
    return parts[0]


def blocked_binary_einsum(inds_c, a: Tensor, b: Tensor) -> Tensor:
    """`binary_einsum(::BackendDagger, inds_c, a, b)` = `Dagger.stage(BinaryEinsum(...))`
    (ext/MuscleDaggerExt/binary_einsum.jl:5-9, 64-119)."""
    from .einsum import binary_einsum, binary_einsum_   # the per-chunk contraction re-enters the dispatch (:60-62)

    ic, ia, ib = list(inds_c), a.inds, b.inds
    # constructor checks, binary_einsum.jl:19-22
    for name, ii in (("ia", ia), ("ib", ib), ("ic", ic)):
        if len(set(ii)) != len(ii):
            raise ArgumentError(f"{name} must have unique indices")
    if not set(ic) <= set(ia) | set(ib):
        raise ArgumentError("ic must be a subset of ia ∪ ib")
    A, B = _to_block_array(a.parent), _to_block_array(b.parent)
    dtype = np.result_type(A.dtype, B.dtype)

    def axis(ii, i):
        return ii.index(i) if i in ii else None

    # output block sizes / grid: from a if the label is there, else from b (:47-58); shared labels must agree
    out_bs, out_grid = [], []
    for i in ic:
        ja, jb = axis(ia, i), axis(ib, i)
        if ja is not None and jb is not None and (A.blocksize[ja] != B.blocksize[jb] or A.shape[ja] != B.shape[jb]):
            raise ArgumentError(f"blocks of batch index {i!r} differ between the operands")
        src, j = (A, ja) if ja is not None else (B, jb)
        out_bs.append(src.blocksize[j])
        out_grid.append(src.chunks.shape[j])
    suminds = [i for i in ia if i in ib and i not in ic]          # setdiff(ia ∪ ib, ic) restricted to shared labels (:82)
    for i in (set(ia) | set(ib)) - set(ic) - set(suminds):
        raise ArgumentError(f"index {i!r} appears in one operand only and not in the output")
    sum_grid = []
    for i in suminds:
        ja, jb = axis(ia, i), axis(ib, i)
        if A.blocksize[ja] != B.blocksize[jb] or A.shape[ja] != B.shape[jb]:
            raise ArgumentError(f"blocks of summed index {i!r} differ between the operands")
        sum_grid.append(A.chunks.shape[ja])

    chunks = np.empty(tuple(out_grid), dtype=object)
    nsum = int(np.prod(sum_grid, dtype=np.int64)) if sum_grid else 1
    for oidx in np.ndindex(*out_grid):
        pos = dict(zip(ic, oidx))
        pairs = []
        for sidx in itertools.product(*[range(g) for g in sum_grid]):   # zip over matching summed blocks (:107-113)
            pos.update(zip(suminds, sidx))
            ca = A.chunks[tuple(pos[i] for i in ia)]
            cb = B.chunks[tuple(pos[i] for i in ib)]
            pairs.append((Tensor(ca, ia), Tensor(cb, ib)))
        on_device = all(t.on_device for p in pairs for t in p)
        if nsum == 1:
            chunks[oidx] = binary_einsum(pairs[0][0], pairs[0][1], out=ic).parent
        elif on_device:
            # every pair straight into its slot of one staging buffer, then one ordered slot sum on the device
            dev = pairs[0][0].parent.device
            slab = int(np.prod(out_bs, dtype=np.int64))
            staging = B200Array((slab * nsum,), dtype, dev)
            esz = np.dtype(dtype).itemsize
            for s, (ta, tb) in enumerate(pairs):
                view = B200Array(tuple(out_bs), dtype, dev, _owner=staging, _ptr=staging.ptr + s * slab * esz)
                binary_einsum_(Tensor(view, ic), ta, tb)
            out = B200Array(tuple(out_bs), dtype, dev)
            h = _lib.Handle.get(dev)
            _lib.check(_lib.lib().mb200_reduce_slots(h.ptr, C.c_void_p(out.ptr), C.c_void_p(staging.ptr), _lib.dtype_enum(dtype),
                                                     slab, nsum))
            chunks[oidx] = out
        else:
            parts = [binary_einsum(ta, tb, out=ic) for ta, tb in pairs]
            parts = [p.to_host().data if p.on_device else p.data for p in parts]
            chunks[oidx] = _lib.fortran(_tree_sum(parts))
    return Tensor(BlockArray(chunks, tuple(out_bs), dtype), ic)
