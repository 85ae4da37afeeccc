"""`unary_einsum(!)` and `hadamard(!)` — the einsum-family neighbours of the hot path, same names,
argument meaning and error behaviour as src/Operations/unary_einsum.jl:26-46 and src/Operations/hadamard.jl:6-38,
with `BackendB200` methods that call libmuscle_b200.so (`mb200_unary_einsum`, `mb200_hadamard`).
Host tensors are staged through HBM (upload, kernel, download); there is no CPU implementation here."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArgumentError
from .backend import Backend, BackendB200, choose_backend
from .tensor import B200Array, Index, Tensor, _as_index_list


def _unique(seq):
    out = []
    for x in seq:
        if x not in out:
            out.append(x)
    return out


def _flatten(inds, extra=()):
    indmap = {}
    for ind in list(inds) + list(extra):
        if ind not in indmap:
            indmap[ind] = len(indmap)
    return indmap


def unary_frontend_inds_y(inds_x, dims=None, out=None):
    """kwargs → inds_y (unary_einsum.jl:26-33): dims defaults to the repeated labels (`nonunique(inds(x))`);
    inds_y = out, or setdiff(inds(x), dims ∩ inds(x))."""
    inds_x = list(inds_x)
    if dims is None:
        dims = [i for i in _unique(inds_x) if inds_x.count(i) > 1]
    elif isinstance(dims, Index):
        dims = [dims]
    dims = _as_index_list(dims)
    inds_sum = [i for i in _unique(dims) if i in inds_x]
    if out is None:
        return [i for i in _unique(inds_x) if i not in inds_sum]
    return _as_index_list(out)


def _b200_unary(inds_y, x: Tensor, y: Tensor | None = None) -> Tensor:
    inds_y = _as_index_list(inds_y)
    for i in inds_y:                      # ext/MuscleOMEinsumExt.jl:32
        if i not in x.inds:
            raise ArgumentError("Output indices must be a subset of input indices")
    indmap = _flatten(x.inds)
    mx, my = [indmap[i] for i in x.inds], [indmap[i] for i in inds_y]
    shape_y = tuple(x.size(i) for i in inds_y)
    host = not x.on_device
    dx = x.data if x.on_device else B200Array.from_host(x.data)
    if y is not None and y.on_device:
        dy = y.data
    else:
        dy = B200Array(shape_y, x.dtype, dx.device)
    if y is not None and (y.shape != shape_y or y.dtype != x.dtype):
        raise _lib.DimensionMismatch(f"size/eltype of y {y.shape}/{y.dtype} does not match {shape_y}/{x.dtype}")
    h = _lib.Handle.get(dx.device)
    _lib.check(_lib.lib().mb200_unary_einsum(
        h.ptr, C.c_void_p(dy.ptr), _lib.dtype_enum(x.dtype), len(my), _lib.i32(my), None,
        C.c_void_p(dx.ptr), _lib.dtype_enum(x.dtype), len(mx), _lib.i32(mx), _lib.i64(dx.shape), None))
    if y is not None:
        if not y.on_device:
            y.data[...] = dy.to_host()
        return y
    return Tensor(dy.to_host() if host else dy, inds_y)


def unary_einsum(*args, dims=None, out=None) -> Tensor:
    """unary_einsum(x; dims=nonunique(inds(x)), out=nothing)        (unary_einsum.jl:26-36)
    unary_einsum(backend, inds_y, x)                               (the per-backend method, :35)"""
    if len(args) == 3 and isinstance(args[0], Backend):
        backend, inds_y, x = args
    elif len(args) == 1 and isinstance(args[0], Tensor):
        x = args[0]
        inds_y = unary_frontend_inds_y(x.inds, dims=dims, out=out)
        backend = choose_backend("unary_einsum", x.parent)
    else:
        raise ArgumentError("unary_einsum(x; dims, out) or unary_einsum(backend, inds_y, x)")
    if isinstance(backend, BackendB200):
        return _b200_unary(inds_y, x)
    raise ArgumentError(f"`unary_einsum` not implemented or not loaded for backend {backend!r}")   # :38-40


def unary_einsum_(*args) -> Tensor:
    """unary_einsum!(y, x) (unary_einsum.jl:42-46) / unary_einsum!(backend, y, x)."""
    if len(args) == 3 and isinstance(args[0], Backend):
        backend, y, x = args
    elif len(args) == 2:
        y, x = args
        backend = choose_backend("unary_einsum!", y.parent, x.parent)
    else:
        raise ArgumentError("unary_einsum!(y, x) or unary_einsum!(backend, y, x)")
    if isinstance(backend, BackendB200):
        return _b200_unary(y.inds, x, y)
    raise ArgumentError(f"`unary_einsum!` not implemented or not loaded for backend {backend!r}")


def _b200_hadamard(a: Tensor, b: Tensor, c: Tensor | None = None) -> Tensor:
    if a.ndim < b.ndim:                   # hadamard.jl:8 `b` must be broadcastable to `a`
        a, b = b, a
    for i in b.inds:                      # hadamard.jl:10
        if i not in a.inds:
            raise ArgumentError("inds(b) ⊆ inds(a) must hold")
    if c is not None and c.inds != a.inds:   # hadamard.jl:28
        raise ArgumentError("inds(c) == inds(a) must hold")
    T = np.result_type(a.dtype, b.dtype)
    indmap = _flatten(a.inds)
    ma, mb = [indmap[i] for i in a.inds], [indmap[i] for i in b.inds]
    host = not a.on_device and not b.on_device and (c is None or not c.on_device)
    dev = next((t.data.device for t in (c, a, b) if t is not None and t.on_device), None)
    da = a.data if a.on_device else B200Array.from_host(a.data, dev)
    db = b.data if b.on_device else B200Array.from_host(b.data, dev)
    if c is not None:
        if c.dtype != T or c.shape != a.shape:
            raise _lib.DimensionMismatch(f"c must have size {a.shape} and eltype {T}")
        dc = c.data if c.on_device else (da if (c is a and a.dtype == T) else B200Array(a.shape, T, da.device))
    else:
        dc = B200Array(a.shape, T, da.device)
    h = _lib.Handle.get(da.device)
    _lib.check(_lib.lib().mb200_hadamard(
        h.ptr, C.c_void_p(dc.ptr), _lib.dtype_enum(T),
        C.c_void_p(da.ptr), _lib.dtype_enum(a.dtype), len(ma), _lib.i32(ma), _lib.i64(da.shape),
        C.c_void_p(db.ptr), _lib.dtype_enum(b.dtype), len(mb), _lib.i32(mb), _lib.i64(db.shape)))
    if c is not None:
        if not c.on_device:
            c.data[...] = dc.to_host()
        return c
    return Tensor(dc.to_host() if host else dc, a.inds)


def hadamard(*args) -> Tensor:
    """hadamard(a, b) (hadamard.jl:6-13): element-wise product, `b` broadcast over the labels it lacks; the result
    carries the labels of the higher-rank operand. hadamard(backend, a, b) is the per-backend method (:12)."""
    if len(args) == 3 and isinstance(args[0], Backend):
        backend, a, b = args
    elif len(args) == 2:
        a, b = args
        if a.ndim < b.ndim:
            a, b = b, a
        for i in b.inds:
            if i not in a.inds:
                raise ArgumentError("inds(b) ⊆ inds(a) must hold")
        backend = choose_backend("hadamard", a.parent, b.parent)
    else:
        raise ArgumentError("hadamard(a, b) or hadamard(backend, a, b)")
    if isinstance(backend, BackendB200):
        return _b200_hadamard(a, b)
    raise ArgumentError(f"`hadamard` not implemented or not loaded for backend {backend!r}")   # :15-17


def hadamard_(*args) -> Tensor:
    """hadamard!(c, a, b) (hadamard.jl:19-34): writes c (inds(c) == inds(a)), returns c; c may be a itself."""
    if len(args) == 4 and isinstance(args[0], Backend):
        backend, c, a, b = args
    elif len(args) == 3:
        c, a, b = args
        backend = choose_backend("hadamard!", c.parent, a.parent, b.parent)
    else:
        raise ArgumentError("hadamard!(c, a, b) or hadamard!(backend, c, a, b)")
    if isinstance(backend, BackendB200):
        return _b200_hadamard(a, b, c)
    raise ArgumentError(f"`hadamard!` not implemented or not loaded for backend {backend!r}")
