"""`binary_einsum` / `binary_einsum!` — same names, argument meaning and error behaviour as
src/Operations/binary_einsum.jl:33-74, with `BackendB200` methods that call libmuscle_b200.so.

The keyword front-end (`dims`, `out` → `inds_c`) runs before the backend is chosen
(binary_einsum.jl:34-41); the backend only ever sees `inds_c`. Labels are flattened to int mode
ids exactly like the reference's cuTENSOR extension (ext/MuscleCUDAExt.jl:24-27) and never cross
the C ABI.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArgumentError
from .backend import Backend, BackendB200, BackendBlocks, choose_backend
from .tensor import B200Array, Index, Tensor, _as_index_list


def _unique(seq):
    out = []
    for x in seq:
        if x not in out:
            out.append(x)
    return out


def frontend_inds_c(inds_a, inds_b, dims=None, out=None):
    """kwargs → inds_c (binary_einsum.jl:33-41):
    inds_sum = dims ∩ inds(a) ∩ inds(b); inds_c = out, or setdiff(inds(a) ∪ inds(b), inds_sum)."""
    if dims is None:
        dims = [i for i in _unique(inds_a) if i in inds_b]
    elif isinstance(dims, Index):
        dims = [dims]
    dims = _as_index_list(dims)
    inds_sum = [i for i in _unique(dims) if i in inds_a and i in inds_b]
    if out is None:
        return [i for i in _unique(list(inds_a) + list(inds_b)) if i not in inds_sum]
    return _as_index_list(out)


def flatten_labels(inds_a, inds_b, inds_c):
    """Index → int mode ids: `indmap = Dict(ind => i for (i, ind) in enumerate(unique(inds_a ∪ inds_b)))`
    (ext/MuscleCUDAExt.jl:24-27). A label of C found in neither operand is an ArgumentError
    (ext/MuscleStridedExt.jl:53)."""
    indmap = {}
    for ind in list(inds_a) + list(inds_b):
        if ind not in indmap:
            indmap[ind] = len(indmap)
    for ind in inds_c:
        if ind not in indmap:
            raise ArgumentError(f"index {ind!r} of the output is found in neither operand")
    return ([indmap[i] for i in inds_a], [indmap[i] for i in inds_b], [indmap[i] for i in inds_c])


def _result_shape(inds_c, a: Tensor, b: Tensor):
    shape = []
    for i in inds_c:
        if i in a.inds:
            shape.append(a.size(i))
        elif i in b.inds:
            shape.append(b.size(i))
        else:
            raise ArgumentError(f"index {i!r} of the output is found in neither operand")
    return tuple(shape)


def _promote(a: Tensor, b: Tensor) -> np.dtype:
    _lib.dtype_enum(a.dtype)
    _lib.dtype_enum(b.dtype)
    return np.result_type(a.dtype, b.dtype)   # Base.promote_eltype (ext/MuscleCUDAExt.jl:16)


def _common_device(*tensors) -> int:
    """The one GPU all device-resident operands live on; operands on different devices are an ArgumentError
    (a kernel launched on one device must never be handed another device's pointers)."""
    devs = {t.data.device for t in tensors if t.on_device}
    if len(devs) > 1:
        raise ArgumentError(f"operands live on different devices {sorted(devs)}; move them to one GPU first")
    return devs.pop()


# Launch-bound regime (every contraction in the reference's own test-suite): the label bookkeeping of a call is a pure
# function of (labels, shapes, eltypes), so it is computed once per signature and the ctypes argument arrays are reused.
# Entry: (T, dtype enum of T, shape_c, mc, ma, mb, ea, eb, len(mc), len(ma), len(mb)) with the ctypes arrays ready to pass.
_SIG_CACHE: dict = {}
_FRONT_CACHE: dict = {}
_SIG_CACHE_CAP = 4096


def _signature(inds_c, a: Tensor, b: Tensor):
    key = (inds_c, a._inds, b._inds, a.data.shape, b.data.shape, a.data.dtype, b.data.dtype)
    sig = _SIG_CACHE.get(key)
    if sig is None:
        ma, mb, mc = flatten_labels(a.inds, b.inds, inds_c)
        T = _promote(a, b)
        shape_c = _result_shape(inds_c, a, b)
        if len(_SIG_CACHE) >= _SIG_CACHE_CAP:
            _SIG_CACHE.clear()
        nbytes = (int(np.prod(shape_c, dtype=np.int64)) if shape_c else 1) * T.itemsize
        sig = _SIG_CACHE[key] = (T, _lib.dtype_enum(T), shape_c, _lib.i32(mc), _lib.i32(ma), _lib.i32(mb),
                                 _lib.i64(a.shape), _lib.i64(b.shape), len(mc), len(ma), len(mb),
                                 _lib.dtype_enum(a.dtype), _lib.dtype_enum(b.dtype), nbytes)
    return sig


def _b200_device_fast(inds_c: tuple, a: Tensor, b: Tensor) -> Tensor:
    """Both operands device-resident: cached bookkeeping, one allocation, one C-ABI call."""
    da, db = a.data, b.data
    if da.device != db.device:
        raise ArgumentError(f"operands live on different devices {sorted({da.device, db.device})}; move them to one GPU first")
    T, eT, shape_c, mc, ma, mb, ea, eb, nc, na, nb, eA, eB, nbytes = _signature(inds_c, a, b)
    dc = B200Array._fast(shape_c, T, da.device, nbytes)
    st = _LIB_CALL[0](dc.handle._h, dc.ptr, eT, nc, mc, None, da.ptr, eA, na, ma, ea, None, db.ptr, eB, nb, mb, eb, None)
    if st:
        _lib.check(st)
    return Tensor._trusted(dc, inds_c)


_LIB_CALL = [None]


def _b200_out_of_place(inds_c, a: Tensor, b: Tensor) -> Tensor:
    """`binary_einsum(::BackendB200, inds_c, a, b)`: allocates C, returns Tensor(C, inds_c)."""
    if a.on_device and b.on_device:
        if _LIB_CALL[0] is None:
            _LIB_CALL[0] = _lib.lib().mb200_binary_einsum
        return _b200_device_fast(tuple(_as_index_list(inds_c)), a, b)
    inds_c = _as_index_list(inds_c)
    ma, mb, mc = flatten_labels(a.inds, b.inds, inds_c)
    T = _promote(a, b)
    L = _lib.lib()
    if not a.on_device and not b.on_device:
        # host-only validation first: argument errors must not depend on a GPU being present
        _lib.plan_describe(_lib.dtype_enum(T), mc, _lib.dtype_enum(a.dtype), ma, a.shape,
                           _lib.dtype_enum(b.dtype), mb, b.shape)
        # host arrays: the library stages through HBM itself (H2D, kernels, D2H, sync)
        h = _lib.Handle.get()
        ha = _lib.fortran(a.data)
        hb = _lib.fortran(b.data)
        # argument validation happens inside the call before any device work; shape needs valid labels
        shape_c = _result_shape(inds_c, a, b)
        hc = np.empty(shape_c, dtype=T, order="F")
        _lib.check(L.mb200_binary_einsum_host(
            h.ptr,
            C.c_void_p(hc.ctypes.data), _lib.dtype_enum(T), len(mc), _lib.i32(mc),
            C.c_void_p(ha.ctypes.data), _lib.dtype_enum(ha.dtype), len(ma), _lib.i32(ma), _lib.i64(ha.shape),
            C.c_void_p(hb.ctypes.data), _lib.dtype_enum(hb.dtype), len(mb), _lib.i32(mb), _lib.i64(hb.shape)))
        return Tensor(hc, inds_c)
    # device path (a host operand of a mixed pair is uploaded first — "hybrid" operands,
    # cf. binary_einsum.jl:23-24)
    dev = _common_device(a, b)
    da = a.data if a.on_device else B200Array.from_host(a.data, dev)
    db = b.data if b.on_device else B200Array.from_host(b.data, dev)
    shape_c = _result_shape(inds_c, a, b)
    dc = B200Array(shape_c, T, dev)
    h = _lib.Handle.get(dev)
    _lib.check(L.mb200_binary_einsum(
        h.ptr,
        C.c_void_p(dc.ptr), _lib.dtype_enum(T), len(mc), _lib.i32(mc), None,
        C.c_void_p(da.ptr), _lib.dtype_enum(da.dtype), len(ma), _lib.i32(ma), _lib.i64(da.shape), None,
        C.c_void_p(db.ptr), _lib.dtype_enum(db.dtype), len(mb), _lib.i32(mb), _lib.i64(db.shape), None))
    return Tensor(dc, inds_c)


def _b200_in_place(c: Tensor, a: Tensor, b: Tensor) -> Tensor:
    """`binary_einsum!(::BackendB200, c, a, b)`: writes parent(c) in inds(c) order, returns c.
    Unlike BackendBase (binary_einsum.jl:108) any order of inds(c) is accepted, as cuTENSOR does."""
    ma, mb, mc = flatten_labels(a.inds, b.inds, c.inds)
    T = _promote(a, b)
    if c.dtype != T:
        raise ArgumentError(f"eltype(c) = {c.dtype} must be promote_eltype(a, b) = {T}")
    if c.shape != _result_shape(c.inds, a, b):
        raise _lib.DimensionMismatch(f"size(c) = {c.shape} does not match the contraction {_result_shape(c.inds, a, b)}")
    L = _lib.lib()
    if c.on_device:
        dev = _common_device(c, a, b)
        da = a.data if a.on_device else B200Array.from_host(a.data, dev)
        db = b.data if b.on_device else B200Array.from_host(b.data, dev)
        h = _lib.Handle.get(dev)
        _lib.check(L.mb200_binary_einsum(
            h.ptr,
            C.c_void_p(c.data.ptr), _lib.dtype_enum(T), len(mc), _lib.i32(mc), None,
            C.c_void_p(da.ptr), _lib.dtype_enum(da.dtype), len(ma), _lib.i32(ma), _lib.i64(da.shape), None,
            C.c_void_p(db.ptr), _lib.dtype_enum(db.dtype), len(mb), _lib.i32(mb), _lib.i64(db.shape), None))
        return c
    if a.on_device or b.on_device:
        raise ArgumentError("binary_einsum!: c on the host needs host operands")
    res = _b200_out_of_place(c.inds, a, b)
    c.data[...] = res.data
    return c


def binary_einsum(*args, dims=None, out=None) -> Tensor:
    """binary_einsum(a, b; dims=∩(inds(a), inds(b)), out=nothing)      (binary_einsum.jl:33-51)
    binary_einsum(backend, inds_c, a, b)                              (the per-backend method, :50)"""
    if len(args) == 4 and isinstance(args[0], Backend):
        backend, inds_c, a, b = args
        if dims is not None or out is not None:
            raise ArgumentError("the backend method takes inds_c, not dims/out")
    elif len(args) == 2:
        a, b = args
        if not isinstance(a, Tensor) or not isinstance(b, Tensor):
            raise ArgumentError("binary_einsum(a::Tensor, b::Tensor; dims, out)")
        # kwargs -> inds_c is a pure function of the labels: cached per (labels, dims, out)
        try:
            fkey = (a._inds, b._inds, None if dims is None else (dims if isinstance(dims, Index) else tuple(dims)),
                    None if out is None else tuple(out))
            inds_c = _FRONT_CACHE.get(fkey)
        except TypeError:
            fkey, inds_c = None, None
        if inds_c is None:
            inds_c = tuple(frontend_inds_c(a.inds, b.inds, dims=dims, out=out))
            if fkey is not None:
                if len(_FRONT_CACHE) >= _SIG_CACHE_CAP:
                    _FRONT_CACHE.clear()
                _FRONT_CACHE[fkey] = inds_c
        backend = choose_backend("binary_einsum", a.data, b.data)
        if type(backend) is BackendB200 and a.on_device and b.on_device:
            if _LIB_CALL[0] is None:
                _LIB_CALL[0] = _lib.lib().mb200_binary_einsum
            return _b200_device_fast(inds_c, a, b)
        inds_c = list(inds_c)
    else:
        raise ArgumentError("binary_einsum(a, b; dims, out) or binary_einsum(backend, inds_c, a, b)")
    if isinstance(backend, BackendB200):
        return _b200_out_of_place(inds_c, a, b)
    if isinstance(backend, BackendBlocks):
        from .blocks import blocked_binary_einsum
        return blocked_binary_einsum(inds_c, a, b)
    # binary_einsum.jl:53-55
    raise ArgumentError(f"`binary_einsum` not implemented or not loaded for backend {backend!r} "
                        "(this package provides BackendB200 only; use with_backend(f, BackendB200()) "
                        "or device-resident tensors)")


def binary_einsum_(*args) -> Tensor:
    """binary_einsum!(c, a, b) (binary_einsum.jl:57-70) / binary_einsum!(backend, c, a, b) (:68)."""
    if len(args) == 4 and isinstance(args[0], Backend):
        backend, c, a, b = args
    elif len(args) == 3:
        c, a, b = args
        backend = choose_backend("binary_einsum!", c.parent, a.parent, b.parent)
    else:
        raise ArgumentError("binary_einsum!(c, a, b) or binary_einsum!(backend, c, a, b)")
    if isinstance(backend, BackendB200):
        return _b200_in_place(c, a, b)
    raise ArgumentError(f"`binary_einsum!` not implemented or not loaded for backend {backend!r}")


binary_einsum_inplace = binary_einsum_
