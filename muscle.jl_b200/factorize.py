"""`tensor_svd_thin` and `simple_update` on the device — the main in-tree caller of the hot path.
Same names, keyword meaning and error behaviour as src/Operations/tensor_svd.jl:62-124, src/Index.jl:41-63
(`factorinds`) and src/Operations/simple_update.jl:15-82. Everything heavy runs in libmuscle_b200.so: K1 permute
(matricise), the hand-written one-sided Jacobi SVD (`mb200_svd_thin`), `binary_einsum` for Θ, `hadamard` for the
absorb step. Only the k singular values are ever touched on the host (normalisation / square roots of a length-k
vector, as the reference does with `normalize!(S)` / `sqrt.(S)`)."""
from __future__ import annotations

import ctypes as C
import itertools

import numpy as np

from . import _lib
from ._lib import ArgumentError
from .backend import Backend, BackendB200, choose_backend
from .einsum import binary_einsum, frontend_inds_c
from .family import hadamard_
from .tensor import B200Array, Index, Tensor, _as_index_list

_gensym = itertools.count()


def factorinds(all_inds, left_inds, right_inds):
    """src/Index.jl:41-63: complete the missing side, check disjointness / membership / non-emptiness."""
    left = [left_inds] if isinstance(left_inds, Index) else _as_index_list(left_inds or [])
    right = [right_inds] if isinstance(right_inds, Index) else _as_index_list(right_inds or [])
    all_inds = list(all_inds)
    if set(left) & set(right):
        raise ArgumentError(f"left ({left}) and right ({right}) indices must be disjoint")
    if not left:
        left = [i for i in all_inds if i not in right]
    elif not right:
        right = [i for i in all_inds if i not in left]
    if not left or not right:
        raise ArgumentError("no right-indices left in factorization")
    if not all(i in all_inds for i in left + right):
        raise ArgumentError(f"indices must be in {all_inds}")
    return left, right


def _real_dtype(dt):
    return np.dtype(np.float32) if np.dtype(dt) in (np.dtype(np.float32), np.dtype(np.complex64)) else np.dtype(np.float64)


def _b200_svd_thin(A: Tensor, inds_u=(), inds_v=(), ind_s=None, tol=0.0, max_sweeps=0):
    if ind_s is None:
        ind_s = Index(("svd", next(_gensym)))
    ind_s = ind_s if isinstance(ind_s, Index) else Index(ind_s)
    inds_u, inds_v = factorinds(A.inds, inds_u, inds_v)
    if set(inds_u) | set(inds_v) != set(A.inds):          # tensor_svd.jl:105
        raise ArgumentError("issetequal(inds_u ∪ inds_v, inds(A)) must hold")
    if ind_s in A.inds:                                    # tensor_svd.jl:106
        raise ArgumentError("ind_s ∉ inds(A) must hold")
    host = not A.on_device
    Ad = A if A.on_device else A.to_device()
    left_sizes = tuple(Ad.size(i) for i in inds_u)
    right_sizes = tuple(Ad.size(i) for i in inds_v)
    rows = int(np.prod(left_sizes, dtype=np.int64))
    cols = int(np.prod(right_sizes, dtype=np.int64))
    k = min(rows, cols)
    order = inds_u + inds_v
    Amat = Ad if Ad.inds == order else Ad.permutedims(order)     # K1 (tensor_svd.jl:111)
    dev = Amat.data.device
    U = B200Array(left_sizes + (k,), A.dtype, dev)
    S = B200Array((k,), _real_dtype(A.dtype), dev)
    Vt = B200Array(right_sizes + (k,), A.dtype, dev)
    h = _lib.Handle.get(dev)
    _lib.check(_lib.lib().mb200_svd_thin(h.ptr, C.c_void_p(U.ptr), C.c_void_p(S.ptr), C.c_void_p(Vt.ptr),
                                         C.c_void_p(Amat.data.ptr), _lib.dtype_enum(A.dtype), rows, cols,
                                         float(tol), int(max_sweeps)))
    out = (Tensor(U, inds_u + [ind_s]), Tensor(S, [ind_s]), Tensor(Vt, inds_v + [ind_s]))
    return tuple(t.to_host() for t in out) if host else out


def svd_last_info(device=None) -> dict:
    """Status of the last `tensor_svd_thin` on this device / thread (synchronises): Jacobi sweeps, converged flag, number of
    null columns completed to an orthonormal basis (rank deficiency)."""
    h = _lib.Handle.get(device)
    a, b, c = C.c_int(), C.c_int(), C.c_int()
    _lib.check(_lib.lib().mb200_svd_last_info(h.ptr, C.byref(a), C.byref(b), C.byref(c)))
    return {"sweeps": int(a.value), "converged": bool(b.value), "completed_columns": int(c.value)}


def tensor_svd_thin(*args, inds_u=(), inds_v=(), ind_s=None, **kwargs):
    """tensor_svd_thin(A; inds_u, inds_v, ind_s) -> U[inds_u..., s], s[s], Vt[inds_v..., s] with
    A = Σ_s U·s·Vt (tensor_svd.jl:62-65, :100-124); tensor_svd_thin(backend, A; ...) is the per-backend method."""
    if len(args) == 2 and isinstance(args[0], Backend):
        backend, A = args
    elif len(args) == 1 and isinstance(args[0], Tensor):
        A = args[0]
        backend = choose_backend("tensor_svd_thin", A.parent)
    else:
        raise ArgumentError("tensor_svd_thin(A; inds_u, inds_v, ind_s)")
    if isinstance(backend, BackendB200):
        return _b200_svd_thin(A, inds_u, inds_v, ind_s, **kwargs)
    raise ArgumentError(f"`tensor_svd_thin` not implemented or not loaded for backend {backend!r}")   # :67-69


def _b200_qr_thin(A: Tensor, inds_q=(), inds_r=(), ind_virtual=None):
    if ind_virtual is None:
        ind_virtual = Index(("qr", next(_gensym)))
    ind_virtual = ind_virtual if isinstance(ind_virtual, Index) else Index(ind_virtual)
    if ind_virtual in A.inds:                                # tensor_qr.jl:60
        raise ArgumentError(f"new virtual bond name ({ind_virtual}) cannot be already be present")
    inds_q, inds_r = factorinds(A.inds, inds_q, inds_r)
    if set(inds_q) | set(inds_r) != set(A.inds):              # tensor_qr.jl:63
        raise ArgumentError("issetequal(inds_q ∪ inds_r, inds(A)) must hold")
    host = not A.on_device
    Ad = A if A.on_device else A.to_device()
    left_sizes = tuple(Ad.size(i) for i in inds_q)
    right_sizes = tuple(Ad.size(i) for i in inds_r)
    rows = int(np.prod(left_sizes, dtype=np.int64))
    cols = int(np.prod(right_sizes, dtype=np.int64))
    k = min(rows, cols)
    order = inds_q + inds_r
    Amat = Ad if Ad.inds == order else Ad.permutedims(order)      # K1 (tensor_qr.jl:68)
    dev = Amat.data.device
    Q = B200Array(left_sizes + (k,), A.dtype, dev)
    R = B200Array((k,) + right_sizes, A.dtype, dev)
    h = _lib.Handle.get(dev)
    _lib.check(_lib.lib().mb200_qr_thin(h.ptr, C.c_void_p(Q.ptr), C.c_void_p(R.ptr), C.c_void_p(Amat.data.ptr),
                                        _lib.dtype_enum(A.dtype), rows, cols))
    out = (Tensor(Q, inds_q + [ind_virtual]), Tensor(R, [ind_virtual] + inds_r))
    return tuple(t.to_host() for t in out) if host else out


def tensor_qr_thin(*args, inds_q=(), inds_r=(), ind_virtual=None):
    """tensor_qr_thin(A; inds_q, inds_r, ind_virtual) -> Q[inds_q..., x], R[x, inds_r...] with A = Q·R and Q isometric
    (tensor_qr.jl:5-79); tensor_qr_thin(backend, A; ...) is the per-backend method."""
    if len(args) == 2 and isinstance(args[0], Backend):
        backend, A = args
    elif len(args) == 1 and isinstance(args[0], Tensor):
        A = args[0]
        backend = choose_backend("tensor_qr_thin", A.parent)
    else:
        raise ArgumentError("tensor_qr_thin(A; inds_q, inds_r, ind_virtual)")
    if isinstance(backend, BackendB200):
        return _b200_qr_thin(A, inds_q, inds_r, ind_virtual)
    raise ArgumentError(f"`tensor_qr_thin` not implemented or not loaded for backend {backend!r}")


def tensor_svd_trunc(*args, inds_u=(), inds_v=(), ind_s=None, threshold=None, maxdim=None, **kwargs):
    """tensor_svd_trunc(A; inds_u, inds_v, ind_s, threshold, maxdim) (tensor_svd.jl:153-201): thin SVD, then keep
    k = min(length(s), maxdim) values and cut at the first one below `threshold * norm(s)`. With the defaults it is
    `tensor_svd_thin`. The cut is a slice of the slowest dimension (no copy); only s visits the host."""
    if len(args) == 2 and isinstance(args[0], Backend):
        backend, A = args
    elif len(args) == 1 and isinstance(args[0], Tensor):
        A = args[0]
        backend = choose_backend("tensor_svd_thin", A.parent)
    else:
        raise ArgumentError("tensor_svd_trunc(A; inds_u, inds_v, ind_s, threshold, maxdim)")
    if not isinstance(backend, BackendB200):
        raise ArgumentError(f"`tensor_svd_trunc` not implemented or not loaded for backend {backend!r}")
    host = not A.on_device
    U, S, Vt = _b200_svd_thin(A.to_device(), inds_u, inds_v, ind_s, **kwargs)
    k = S.shape[0]
    if maxdim is not None:                                    # tensor_svd.jl:181-183
        k = min(k, int(maxdim))
    if threshold is not None:                                 # :185-188
        s = S.to_host().data
        cut = float(np.linalg.norm(s)) * float(threshold)
        below = np.nonzero(s[:k] < cut)[0]
        if below.size:
            k = int(below[0]) + 1                             # Julia's findfirst is 1-based: keep = 1:findfirst(...)
    out = (_slice_last(U, k), _slice_last(S, k), _slice_last(Vt, k))
    return tuple(t.to_host() for t in out) if host else out


def _slice_last(t: Tensor, n: int) -> Tensor:
    """`view(t, ind_s => 1:n)` when ind_s is the last (slowest) dimension: the leading n slabs of the same buffer."""
    if n >= t.shape[-1]:
        return t
    d = t.data
    sub = B200Array(d.shape[:-1] + (n,), d.dtype, d.device, _owner=d, _ptr=d.ptr)
    return Tensor(sub, t.inds)


class AbsorbBehavior:
    """simple_update.jl:7-12."""


class DontAbsorb(AbsorbBehavior):
    pass


class AbsorbU(AbsorbBehavior):
    pass


class AbsorbV(AbsorbBehavior):
    pass


class AbsorbEqually(AbsorbBehavior):
    pass


def _b200_simple_update(A, ind_physical_a, B, ind_physical_b, ind_bond_ab, G, ind_physical_g_a, ind_physical_g_b,
                        normalize=False, absorb=None, maxdim=None):
    absorb = absorb if absorb is not None else DontAbsorb()
    ix = lambda i: i if isinstance(i, Index) else Index(i)
    ind_physical_a, ind_physical_b, ind_bond_ab = ix(ind_physical_a), ix(ind_physical_b), ix(ind_bond_ab)
    ind_physical_g_a, ind_physical_g_b = ix(ind_physical_g_a), ix(ind_physical_g_b)
    host = not (A.on_device or B.on_device or G.on_device)
    A, B, G = A.to_device(), B.to_device(), G.to_device()
    # Θ = binary_einsum(binary_einsum(A, B; dims=[bond]), G; dims=[phys_a, phys_b])          (simple_update.jl:51)
    # The second contraction is asked for Θ directly in the layout the factorisation wants, [inds_u; inds_v] (:54-57), so the
    # permuting epilogue of the GEMM does the matricisation and tensor_svd_thin's permutedims (tensor_svd.jl:111) is a no-op:
    # no K1 pass between the contraction and the SVD.
    inds_u = [i for i in A.inds if i != ind_bond_ab]                                          # :54-56
    inds_v = [i for i in B.inds if i != ind_bond_ab]
    ren = {ind_physical_g_a: ind_physical_a, ind_physical_g_b: ind_physical_b}
    back = {v: k for k, v in ren.items()}
    ab = binary_einsum(A, B, dims=[ind_bond_ab])
    default = frontend_inds_c(ab.inds, G.inds, dims=[ind_physical_a, ind_physical_b])
    want = [back.get(i, i) for i in inds_u + inds_v]
    out = want if sorted(map(repr, want)) == sorted(map(repr, default)) else None   # unusual gates keep the front-end order
    theta = binary_einsum(ab, G, dims=[ind_physical_a, ind_physical_b], out=out)
    # replace(Θ, g_a => a, g_b => b)                                                          (:52)
    theta = Tensor(theta.data, [ren.get(i, i) for i in theta.inds])
    U, S, V = _b200_svd_thin(theta, inds_u=inds_u, inds_v=inds_v, ind_s=ind_bond_ab)          # :57
    if maxdim is not None:                                                                    # :61-65
        n = min(int(maxdim), S.shape[0])
        U, S, V = _slice_last(U, n), _slice_last(S, n), _slice_last(V, n)
    if normalize or isinstance(absorb, AbsorbEqually):
        s = S.to_host().data.copy()        # k numbers
        if normalize:                                                                         # :67
            s = s / np.linalg.norm(s)
            S.data.copy_from_host(s.astype(S.dtype))
    if isinstance(absorb, DontAbsorb):                                                        # :69-80
        res = (U, S, V)
    else:
        if isinstance(absorb, AbsorbU):
            U = hadamard_(U, U, S)
        elif isinstance(absorb, AbsorbV):
            V = hadamard_(V, V, S)
        else:
            sq = Tensor(B200Array.from_host(np.sqrt(s).astype(S.dtype), S.data.device), S.inds)
            U = hadamard_(U, U, sq)
            V = hadamard_(V, V, sq)
        res = (U, V)
    return tuple(t.to_host() for t in res) if host else res


def simple_update(*args, normalize=False, absorb=None, maxdim=None, atol=0.0, rtol=0.0):
    """simple_update(A, ind_physical_a, B, ind_physical_b, ind_bond_ab, G, ind_physical_g_a, ind_physical_g_b;
    normalize, absorb, maxdim) (simple_update.jl:15-82): contract the two sites and the gate, split with a thin SVD
    along the bond, truncate to `maxdim`, optionally normalise / absorb the singular values.
    Returns (U, S, V) for DontAbsorb, (U, V) otherwise."""
    if len(args) == 9 and isinstance(args[0], Backend):
        backend, rest = args[0], args[1:]
    elif len(args) == 8:
        rest = args
        backend = choose_backend("simple_update", rest[0].parent, rest[2].parent, rest[5].parent)
    else:
        raise ArgumentError("simple_update(A, ind_physical_a, B, ind_physical_b, ind_bond_ab, G, ind_physical_g_a, ind_physical_g_b)")
    if isinstance(backend, BackendB200):
        return _b200_simple_update(*rest, normalize=normalize, absorb=absorb, maxdim=maxdim)
    raise ArgumentError(f"`simple_update` not implemented or not loaded for backend {backend!r}")
