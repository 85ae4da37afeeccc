"""TEST INFRASTRUCTURE ONLY — CPU restatement of Muscle.jl v0.3.20 `binary_einsum`.

Julia is not installed in the build image, so the reference cannot be executed; this is a
NumPy (OpenBLAS) restatement of the reference algorithm, column-major like Julia arrays.
It is pinned by the reference's own known-answer tests (tests/test_oracle.py::test_reference_battery_on_oracle
replays test/unit/operations/binary_einsum.jl and the OMEinsum/cuTENSOR integration batteries).
Beyond that tiny known-answer set the reference holds no golden vectors for random data
("parity unpinned" by the reference itself for the large configs).

Arrays are numpy arrays whose `.shape` is the Julia `size`; element (i1,..,iN) is the same
logical element as in Julia. Column-major semantics only matter for `reshape`, which is always
done with order='F' here.

Each function cites the reference lines it follows (paths relative to /root/reference).
"""
from __future__ import annotations

import numpy as np


class ArgumentError(ValueError):
    """Julia `ArgumentError` (raised by `@argcheck`, src/Operations/binary_einsum.jl:82-83)."""


class DimensionMismatch(ValueError):
    """Julia `DimensionMismatch` (src/Tensor.jl:23)."""


def _unique(seq):
    out = []
    for x in seq:
        if x not in out:
            out.append(x)
    return out


def _intersect(*seqs):
    """Julia `∩`: order of the first argument, unique."""
    first, rest = seqs[0], seqs[1:]
    return [x for x in _unique(first) if all(x in s for s in rest)]


def _union(a, b):
    return _unique(list(a) + list(b))


def _setdiff(a, b):
    return [x for x in _unique(a) if x not in b]


def _symdiff(a, b):
    return [x for x in _unique(a) if x not in b] + [x for x in _unique(b) if x not in a]


def _fortran(x: np.ndarray) -> np.ndarray:
    """Column-major copy that keeps 0-dim arrays 0-dim (np.asfortranarray promotes them to 1-d)."""
    return x if x.ndim == 0 else np.asfortranarray(x)


def check_tensor(data: np.ndarray, inds) -> None:
    """Tensor constructor checks — src/Tensor.jl:15-24."""
    if len(inds) != data.ndim:
        raise ArgumentError(f"ndims(data) [{data.ndim}] must be equal to length(inds) [{len(inds)}]")
    for i in _unique(inds):
        sizes = {data.shape[d] for d, j in enumerate(inds) if j == i}
        if len(sizes) > 1:
            raise DimensionMismatch("nonuniform size of repeated indices")


def permutedims(data: np.ndarray, inds, new_inds):
    """`permutedims(t::Tensor, perm::Vector{Index})` — src/Tensor.jl:316-319 → :302-308.

    Label → first position (`findfirst`), then `Base.permutedims` (always a fresh copy).
    0-dim shortcut — src/Tensor.jl:311-312.
    """
    if data.ndim == 0:
        return data, list(inds)
    perm = [list(inds).index(i) for i in new_inds]
    return _fortran(np.transpose(data, perm)), [inds[p] for p in perm]


def frontend_inds_c(inds_a, inds_b, dims=None, out=None):
    """Front-end kwargs → `inds_c` — src/Operations/binary_einsum.jl:33-41.

    inds_sum = dims ∩ inds(a) ∩ inds(b); inds_c = out, or setdiff(inds(a) ∪ inds(b), inds_sum).
    """
    if dims is None:
        dims = _intersect(inds_a, inds_b)
    inds_sum = _intersect(dims, inds_a, inds_b)
    if out is None:
        return _setdiff(_union(inds_a, inds_b), inds_sum)
    return list(out)


def binary_einsum_base(inds_c, a: np.ndarray, inds_a, b: np.ndarray, inds_b):
    """`binary_einsum(::BackendBase, inds_c, a, b)` — src/Operations/binary_einsum.jl:76-96 (TTGT)."""
    check_tensor(a, inds_a)
    check_tensor(b, inds_b)
    inds_a, inds_b, inds_c = list(inds_a), list(inds_b), list(inds_c)
    inds_contract = _intersect(inds_a, inds_b)            # :77  (A's order)
    inds_left = _setdiff(inds_a, inds_contract)           # :78
    inds_right = _setdiff(inds_b, inds_contract)          # :79
    # :82-83  can't deal with hyperindices
    if any(i in inds_contract for i in inds_c) or set(inds_c) != set(_symdiff(inds_a, inds_b)):
        raise ArgumentError("`BackendBase` can't deal with hyperindices.")
    sizes_left = [a.shape[inds_a.index(i)] for i in inds_left]        # :85
    sizes_right = [b.shape[inds_b.index(i)] for i in inds_right]      # :86
    sizes_contract = [a.shape[inds_a.index(i)] for i in inds_contract]  # :87
    M = int(np.prod(sizes_left, dtype=np.int64)) if sizes_left else 1
    N = int(np.prod(sizes_right, dtype=np.int64)) if sizes_right else 1
    K = int(np.prod(sizes_contract, dtype=np.int64)) if sizes_contract else 1
    a_p, _ = permutedims(a, inds_a, inds_left + inds_contract)        # :89
    b_p, _ = permutedims(b, inds_b, inds_contract + inds_right)       # :90
    a_mat = np.reshape(a_p, (M, K), order="F")
    b_mat = np.reshape(b_p, (K, N), order="F")
    c_mat = a_mat @ b_mat                                             # :92  (BLAS gemm)
    c = np.reshape(c_mat, tuple(sizes_left) + tuple(sizes_right), order="F")  # :94
    c, _ = permutedims(c, inds_left + inds_right, inds_c)             # :95
    return _fortran(np.asarray(c))


def binary_einsum_base_inplace(c: np.ndarray, inds_c, a, inds_a, b, inds_b):
    """`binary_einsum!(::BackendBase, c, a, b)` — src/Operations/binary_einsum.jl:98-121.

    Requires inds(c) == [left; right] exactly (:108, no output permutation).
    """
    inds_a, inds_b, inds_c = list(inds_a), list(inds_b), list(inds_c)
    inds_contract = _intersect(inds_a, inds_b)
    inds_left = _setdiff(inds_a, inds_contract)
    inds_right = _setdiff(inds_b, inds_contract)
    if any(i in inds_contract for i in inds_c) or set(inds_c) != set(_symdiff(inds_a, inds_b)):
        raise ArgumentError("`BackendBase` can't deal with hyperindices.")
    if inds_c != inds_left + inds_right:
        raise ArgumentError("inds(c) == [inds_left; inds_right] must hold")
    res = binary_einsum_base(inds_c, a, inds_a, b, inds_b)
    if res.shape != c.shape:
        raise DimensionMismatch(f"output shape {c.shape} != {res.shape}")
    c[...] = res
    return c


def binary_einsum_general(inds_c, a: np.ndarray, inds_a, b: np.ndarray, inds_b):
    """General pairwise einsum incl. hyperindices (batch labels present in A, B and C).

    Semantics of `OMEinsum.einsum!((ia, ib), ic, (A, B), C, true, false, size_dict)`
    — ext/MuscleOMEinsumExt.jl:40-59 — and of `cuTENSOR.contract!(α=1, β=0)`
    — ext/MuscleCUDAExt.jl:22-41; the explicit loop nest the reference tests compare against is
    test/integration/omeinsum.jl:225-236 / test/integration/cuda.jl:169-180:
        C[inds_c] = Σ_{labels not in inds_c} A[inds_a] * B[inds_b].
    """
    check_tensor(a, inds_a)
    check_tensor(b, inds_b)
    labels = _union(inds_a, inds_b)
    for i in inds_c:
        if i not in labels:
            raise ArgumentError(f"index {i!r} of the output found in neither operand")
    for i in labels:
        ea = [a.shape[d] for d, j in enumerate(inds_a) if j == i]
        eb = [b.shape[d] for d, j in enumerate(inds_b) if j == i]
        if ea and eb and ea[0] != eb[0]:
            raise DimensionMismatch(f"extent mismatch for index {i!r}: {ea[0]} vs {eb[0]}")
    m = {lab: k for k, lab in enumerate(labels)}
    T = np.result_type(a.dtype, b.dtype)
    c = np.einsum(a.astype(T, copy=False), [m[i] for i in inds_a],
                  b.astype(T, copy=False), [m[i] for i in inds_b],
                  [m[i] for i in inds_c], optimize=True)
    return _fortran(np.asarray(c))


def binary_einsum(a, inds_a, b, inds_b, dims=None, out=None, general=True):
    """Front-end + backend. `general=False` is exactly Muscle's default host path
    (front-end :33-51 then BackendBase :76-96, which throws on hyperindices);
    `general=True` falls to the OMEinsum/cuTENSOR semantics when hyperindices are present
    (that is what a GPU backend of the reference accepts, test/integration/cuda.jl:126-144)."""
    inds_c = frontend_inds_c(inds_a, inds_b, dims=dims, out=out)
    if general:
        try:
            return binary_einsum_base(inds_c, a, inds_a, b, inds_b), inds_c
        except ArgumentError:
            return binary_einsum_general(inds_c, a, inds_a, b, inds_b), inds_c
    return binary_einsum_base(inds_c, a, inds_a, b, inds_b), inds_c


def rel_frobenius(x, ref) -> float:
    """‖x − ref‖_F / ‖ref‖_F — the criterion of Julia's `isapprox` on arrays and of
    the north-star tolerance (≤1e-12 ComplexF64/Float64, ≤1e-5 ComplexF32/Float32)."""
    x = np.asarray(x)
    ref = np.asarray(ref)
    den = float(np.linalg.norm(ref.ravel()))
    num = float(np.linalg.norm((x.astype(ref.dtype, copy=False) - ref).ravel()))
    return num / den if den > 0 else num
