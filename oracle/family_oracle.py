"""TEST INFRASTRUCTURE ONLY — CPU restatement of Muscle.jl v0.3.20 `unary_einsum` and `hadamard`
(the einsum-family neighbours of the hot path). NumPy, column-major semantics.

Pinned by the reference's own known-answer tests: test/integration/omeinsum.jl:6-100 (unary_einsum: axis sum,
diagonal, trace) and test/unit/operations/hadamard.jl:4-139 (scalar / vector / tensor broadcast), transcribed in
tests/test_family.py. The arithmetic of `unary_einsum` lives in the third-party OMEinsum.jl (compat "0.8, 0.9",
Project.toml; no Manifest is checked in, so the version is unpinned); its published semantics are plain Einstein
summation, restated here as diagonal extraction + axis sums + permutation.
"""
from __future__ import annotations

import numpy as np

from .muscle_oracle import ArgumentError, DimensionMismatch, _fortran, _intersect, _setdiff, _unique, check_tensor


def _nonunique(seq):
    """Muscle's `nonunique`: the labels that occur more than once, in order of first appearance."""
    seq = list(seq)
    return [x for x in _unique(seq) if seq.count(x) > 1]


def unary_frontend_inds_y(inds_x, dims=None, out=None):
    """kwargs → inds_y — src/Operations/unary_einsum.jl:26-33:
    inds_sum = dims ∩ inds(x) (dims defaults to the repeated labels); inds_y = out, or setdiff(inds(x), inds_sum)."""
    if dims is None:
        dims = _nonunique(inds_x)
    inds_sum = _intersect(list(dims), list(inds_x))
    if out is None:
        return _setdiff(list(inds_x), inds_sum)
    return list(out)


def unary_einsum_general(inds_y, x: np.ndarray, inds_x):
    """`unary_einsum!(::BackendOMEinsum, y, x)` — ext/MuscleOMEinsumExt.jl:31-38:
    `@argcheck inds(y) ⊆ inds(x)`, then `einsum!((inds(x),), inds(y), (x,), y, true, false, size_dict)`, i.e.
    y[inds_y] = Σ_{labels of x not in y} x[inds_x] with repeated labels of x tied together (diagonal)."""
    inds_x, inds_y = list(inds_x), list(inds_y)
    check_tensor(x, inds_x)
    for i in inds_y:
        if i not in inds_x:
            raise ArgumentError("Output indices must be a subset of input indices")
    if len(_unique(inds_y)) != len(inds_y):
        raise ArgumentError("repeated output indices are not supported")
    # 1. diagonal of every repeated label: keep its first position
    cur, cur_inds = x, list(inds_x)
    for lab in _nonunique(inds_x):
        pos = [d for d, j in enumerate(cur_inds) if j == lab]
        while len(pos) > 1:
            d0, d1 = pos[0], pos[1]
            diag = np.diagonal(cur, axis1=d0, axis2=d1)           # diagonal axis goes last
            cur = np.moveaxis(diag, -1, d0)                        # put it back at the first position
            cur_inds = [j for d, j in enumerate(cur_inds) if d != d1]
            pos = [d for d, j in enumerate(cur_inds) if j == lab]
    # 2. sum over the labels missing from y
    axes = tuple(d for d, j in enumerate(cur_inds) if j not in inds_y)
    if axes:
        cur = np.sum(cur, axis=axes)
        cur_inds = [j for j in cur_inds if j in inds_y]
    # 3. permute to y's label order
    cur = np.asarray(cur)
    if cur.ndim:
        cur = np.transpose(cur, [cur_inds.index(i) for i in inds_y])
    return _fortran(np.array(cur, dtype=x.dtype))


def unary_einsum(x, inds_x, dims=None, out=None):
    """Front-end + backend (unary_einsum.jl:26-36)."""
    inds_y = unary_frontend_inds_y(inds_x, dims=dims, out=out)
    return unary_einsum_general(inds_y, x, inds_x), inds_y


def hadamard_base(a: np.ndarray, inds_a, b: np.ndarray, inds_b):
    """`hadamard(::BackendBase, a, b)` → `hadamard!(::BackendBase, c, a, b)` — src/Operations/hadamard.jl:42-77.
    Returns (c, inds_c); c has the labels of the higher-rank operand."""
    inds_a, inds_b = list(inds_a), list(inds_b)
    if a.ndim < b.ndim:                                        # :44 / :52  `b` must be broadcastable to `a`
        return hadamard_base(b, inds_b, a, inds_a)
    check_tensor(a, inds_a)
    check_tensor(b, inds_b)
    for i in inds_b:                                           # :10  @argcheck inds(b) ⊆ inds(a)
        if i not in inds_a:
            raise ArgumentError("inds(b) ⊆ inds(a) must hold")
    T = np.result_type(a.dtype, b.dtype)
    if b.ndim == 0:                                            # :57-60 tensor-scalar
        return _fortran(np.asarray(a * b, dtype=T)), inds_a
    shape_b_bcast = [1] * a.ndim                               # :66-71
    for d, ind in enumerate(inds_a):
        if ind in inds_b:
            if a.shape[d] != b.shape[inds_b.index(ind)]:
                raise DimensionMismatch(f"extent mismatch for index {ind!r}")
            shape_b_bcast[d] = b.shape[inds_b.index(ind)]
    data_b = b
    if b.ndim > 1:                                             # :74-77 permute b into a's label order
        order = [i for i in inds_a if i in inds_b]
        data_b = np.transpose(b, [inds_b.index(i) for i in order])
    data_b = np.reshape(np.asfortranarray(data_b), shape_b_bcast, order="F")   # :80
    return _fortran(np.asarray(a * data_b, dtype=T)), inds_a   # :81


def contract_path_oracle(arrays, inds_list, out, path):
    """n-ary contraction as the reference's callers do it: a fold of pairwise `binary_einsum`s along a path
    (src/Operations/simple_update.jl:51-80, test/integration/reactant.jl:111). A label is summed at the step where
    it is no longer needed by any other live tensor nor by `out`; labels dangling in one operand are summed first
    (`unary_einsum`), labels shared by more tensors stay as hyperindices. Returns the result in `out` order."""
    from .muscle_oracle import binary_einsum_general
    live = {i: (np.asarray(a), list(ix)) for i, (a, ix) in enumerate(zip(arrays, inds_list))}
    nxt = len(arrays)
    out = list(out)
    for (a, b) in path:
        xa, ia = live.pop(a)
        xb, ib = live.pop(b)
        needed = set(out)
        for _, ix in live.values():
            needed.update(ix)
        ops = []
        for (x, ix, other) in ((xa, ia, ib), (xb, ib, ia)):
            keep = [l for l in ix if l in other or l in needed]
            if len(keep) != len(ix):
                x = unary_einsum_general(keep, x, ix)
                ix = keep
            ops.append((x, ix))
        (xa, ia), (xb, ib) = ops
        ic = [l for l in dict.fromkeys(ia + ib) if l in needed]
        live[nxt] = (binary_einsum_general(ic, xa, ia, xb, ib), ic)
        nxt += 1
    (x, ix), = live.values()
    return unary_einsum_general(out, x, ix)


def tensor_svd_thin_base(a: np.ndarray, inds_a, inds_u, inds_v):
    """`tensor_svd_thin(::BackendBase, A; inds_u, inds_v)` — src/Operations/tensor_svd.jl:100-124: permutedims to
    [inds_u..., inds_v...], reshape to a matrix, LAPACK thin SVD, tensorify: U[inds_u..., s], s, Vt = conj(V)[inds_v..., s]
    (numpy returns V^H, so Vt = (V^H)^T)."""
    inds_a, inds_u, inds_v = list(inds_a), list(inds_u), list(inds_v)
    left = tuple(a.shape[inds_a.index(i)] for i in inds_u)
    right = tuple(a.shape[inds_a.index(i)] for i in inds_v)
    amat = np.reshape(np.asfortranarray(np.transpose(a, [inds_a.index(i) for i in inds_u + inds_v])),
                      (int(np.prod(left)), int(np.prod(right))), order="F")
    u, s, vh = np.linalg.svd(amat, full_matrices=False)
    k = s.shape[0]
    U = np.reshape(np.asfortranarray(u), left + (k,), order="F")
    Vt = np.reshape(np.asfortranarray(vh.T), right + (k,), order="F")
    return _fortran(U), s, _fortran(Vt)


def simple_update_theta(a, inds_a, b, inds_b, g, inds_g, ind_pa, ind_pb, ind_bond, ind_ga, ind_gb):
    """Θ of `simple_update` — src/Operations/simple_update.jl:51-52: contract A·B over the bond, then with G over the
    physical indices, and rename G's output physical indices to the sites' ones. Returns (Θ, inds)."""
    from .muscle_oracle import binary_einsum
    ab, iab = binary_einsum(a, list(inds_a), b, list(inds_b), dims=[ind_bond])
    th, ith = binary_einsum(ab, iab, g, list(inds_g), dims=[ind_pa, ind_pb])
    ren = {ind_ga: ind_pa, ind_gb: ind_pb}
    return th, [ren.get(i, i) for i in ith]


def dagger_stage_oracle(inds_c, a: np.ndarray, inds_a, blocks_a, b: np.ndarray, inds_b, blocks_b):
    """Restatement of `Dagger.stage(::BinaryEinsum)` (ext/MuscleDaggerExt/binary_einsum.jl:64-119) on plain numpy blocks:
    output block sizes from a, else b (:47-58); grid = size ÷ blocksize (:68-69); per output block the add-tree reduction
    (`treereduce(AddComputeOp)`, :107-115) over the summed-index blocks of one chunk contraction each
    (`task_binary_einsum`, :60-62). Returns (dense result, block sizes of the result, list of chunk shapes).
    Pinned on the reference's own test (test/integration/dagger.jl:12-31: 2 x 2 Float64 in 1 x 1 blocks,
    `collect(block_c) ≈ c`, every chunk of size (1, 1))."""
    from .muscle_oracle import binary_einsum_general
    ia, ib, ic = list(inds_a), list(inds_b), list(inds_c)
    for ii in (ia, ib, ic):                                  # :19-21
        if len(set(ii)) != len(ii):
            raise ArgumentError("indices must be unique")
    if not set(ic) <= set(ia) | set(ib):                     # :22
        raise ArgumentError("ic must be a subset of ia ∪ ib")
    bs = {}
    for ii, blk, arr in ((ib, blocks_b, b), (ia, blocks_a, a)):   # a's block size wins for shared labels (:48-55)
        for j, i in enumerate(ii):
            if arr.shape[j] % blk[j]:
                raise ArgumentError("extent is not a multiple of the block size")
            bs[i] = blk[j]
    size = {i: a.shape[j] for j, i in enumerate(ia)}
    size.update({i: b.shape[j] for j, i in enumerate(ib) if i not in size})
    suminds = [i for i in ia if i in ib and i not in ic]      # :82
    out = np.zeros([size[i] for i in ic], dtype=np.result_type(a.dtype, b.dtype), order="F")
    chunk_shapes = []
    grid = [size[i] // bs[i] for i in ic]
    for oidx in np.ndindex(*grid):
        pos = dict(zip(ic, oidx))
        parts = []
        for sidx in np.ndindex(*[size[i] // bs[i] for i in suminds]):
            pos.update(zip(suminds, sidx))
            ca = a[tuple(slice(pos[i] * bs[i], (pos[i] + 1) * bs[i]) for i in ia)]
            cb = b[tuple(slice(pos[i] * bs[i], (pos[i] + 1) * bs[i]) for i in ib)]
            parts.append(binary_einsum_general(ic, _fortran(ca), ia, _fortran(cb), ib))
        while len(parts) > 1:                                 # treereduce: pairwise add tree
            nxt = [parts[k] + parts[k + 1] for k in range(0, len(parts) - 1, 2)]
            if len(parts) % 2:
                nxt.append(parts[-1])
            parts = nxt
        out[tuple(slice(o * bs[i], (o + 1) * bs[i]) for o, i in zip(oidx, ic))] = parts[0]
        chunk_shapes.append(parts[0].shape)
    return out, tuple(bs[i] for i in ic), chunk_shapes
