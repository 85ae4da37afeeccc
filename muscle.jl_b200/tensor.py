"""Host-side mirror of Muscle's data model, restricted to what `binary_einsum` needs:
`Index` (src/Index.jl:3-5), `Tensor` (src/Tensor.jl:11-32) with `inds` / `parent` / `dim` / `size`,
and `B200Array`, the device-resident counterpart of a Julia `Array` (dense, column-major, complex
interleaved) that selects `BackendB200` through the Domain rules.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArgumentError, DimensionMismatch


class Index:
    """`Index{T}`: an opaque label; equality is tag equality (src/Index.jl:3-5)."""
    __slots__ = ("tag",)

    def __init__(self, tag):
        self.tag = tag.tag if isinstance(tag, Index) else tag

    def __eq__(self, other):
        return isinstance(other, Index) and self.tag == other.tag

    def __hash__(self):
        return hash(("Index", self.tag))

    def __repr__(self):
        return f"index<{self.tag}>"


def _as_index_list(inds):
    return [i if isinstance(i, Index) else Index(i) for i in inds]


def findperm(frm, to):
    """`findperm` tolerant of repeated labels (src/Index.jl:20-34)."""
    frm, to = list(frm), list(to)
    if sorted(map(repr, frm)) != sorted(map(repr, to)):
        raise AssertionError("issetequal(from, to)")
    used = [False] * len(to)
    perm = []
    for ind in frm:
        for k, t in enumerate(to):
            if not used[k] and t == ind:
                used[k] = True
                perm.append(k)
                break
    return perm


_TORCH_DEV: dict = {}   # device index -> (torch module, torch.device, torch.uint8) or False


def _torch_device(index):
    try:
        import torch
        ok = torch.cuda.is_available()
    except ImportError:
        ok = False
    _TORCH_DEV[index] = (torch, torch.device("cuda", index), torch.uint8) if ok else False
    return _TORCH_DEV[index]


class B200Array:
    """Dense column-major array in B200 HBM. Memory comes from torch's caching allocator when torch
    sees a GPU (so it is ordered with torch streams), else from mb200_malloc."""

    def __init__(self, shape, dtype, device=None, _owner=None, _ptr=None):
        self.shape = tuple(int(s) for s in shape)
        self.dtype = np.dtype(dtype)
        _lib.dtype_enum(self.dtype)
        self.handle = _lib.Handle.get(device)
        self.device = self.handle.device
        self.nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize if self.shape else self.dtype.itemsize
        self._owner = _owner
        self._own_malloc = False
        if _ptr is not None:
            self.ptr = int(_ptr)
        else:
            self.ptr = self._alloc(max(self.nbytes, 1))

    @classmethod
    def _fast(cls, shape, dtype, device, nbytes):
        """Trusted constructor of the launch-bound fast path: `shape` a tuple of ints, `dtype` a supported np.dtype, `nbytes`
        precomputed; memory from torch's caching allocator (no dtype / shape re-validation)."""
        self = object.__new__(cls)
        self.shape, self.dtype, self.nbytes = shape, dtype, nbytes
        h = self.handle = _lib.Handle.get(device)
        self.device = h.device
        self._own_malloc = False
        tv = _TORCH_DEV.get(h.device)
        if tv is None:
            tv = _torch_device(h.device)
        if tv is False:
            self._owner = None
            self.ptr = self._alloc(max(nbytes, 1))
        else:
            torch, dev, u8 = tv
            self._owner = torch.empty(max(nbytes, 1), dtype=u8, device=dev)
            self.ptr = self._owner.data_ptr()
        return self

    def _alloc(self, nbytes):
        try:
            import torch
            if torch.cuda.is_available():
                self._owner = torch.empty(nbytes, dtype=torch.uint8, device=f"cuda:{self.device}")
                return int(self._owner.data_ptr())
        except ImportError:
            pass
        p = C.c_void_p()
        _lib.check(_lib.lib().mb200_malloc(self.handle.ptr, C.byref(p), nbytes))
        self._own_malloc = True
        return int(p.value)

    def __del__(self):
        if getattr(self, "_own_malloc", False):
            try:
                _lib.lib().mb200_free(self.handle.ptr, C.c_void_p(self.ptr))
            except Exception:
                pass

    @property
    def ndim(self):
        return len(self.shape)

    @property
    def size(self):
        return int(np.prod(self.shape, dtype=np.int64)) if self.shape else 1

    @classmethod
    def from_host(cls, array, device=None, non_blocking=False):
        """Upload on the current stream. non_blocking=True returns without synchronising: the source must be
        pinned host memory and must stay alive and unmodified until the stream reaches the copy."""
        a = _lib.fortran(array)
        out = cls(a.shape, a.dtype, device)
        out.copy_from_host(a, non_blocking=non_blocking)
        return out

    @classmethod
    def from_torch(cls, t, shape, dtype):
        """Alias a contiguous CUDA torch tensor as a column-major array of `shape`/`dtype` (no copy; the
        torch tensor is kept alive as the owner). Used for device-generated synthetic data."""
        dtype = np.dtype(dtype)
        n = int(np.prod(shape, dtype=np.int64)) if len(shape) else 1
        if not t.is_cuda or not t.is_contiguous() or t.numel() * t.element_size() < n * dtype.itemsize:
            raise ArgumentError("from_torch needs a contiguous CUDA tensor at least as large as the array")
        import torch
        owner = t.view(torch.uint8).reshape(-1) if t.dtype != torch.uint8 else t.reshape(-1)
        return cls(shape, dtype, t.device.index, _owner=owner, _ptr=t.data_ptr())

    def copy_from_host(self, a, non_blocking=False):
        a = _lib.fortran(a)
        if a.shape != self.shape or a.dtype != self.dtype:
            raise DimensionMismatch(f"copy_from_host: {a.shape}/{a.dtype} into {self.shape}/{self.dtype}")
        h = _lib.Handle.get(self.device)
        if a.nbytes:
            _lib.check(_lib.lib().mb200_memcpy_h2d(h.ptr, C.c_void_p(self.ptr), C.c_void_p(a.ctypes.data), a.nbytes))
            if not non_blocking:
                h.synchronize()  # the numpy source may be pageable and may die after return

    def to_host(self, out=None, non_blocking=False):
        """Download on the current stream. With non_blocking=True `out` must be pinned and is only valid after the
        stream has been synchronised by the caller."""
        if out is None:
            out = np.empty(self.shape, dtype=self.dtype, order="F")
        h = _lib.Handle.get(self.device)
        if out.nbytes:
            _lib.check(_lib.lib().mb200_memcpy_d2h(h.ptr, C.c_void_p(out.ctypes.data), C.c_void_p(self.ptr), out.nbytes))
        if not non_blocking:
            h.synchronize()
        return out

    def __repr__(self):
        return f"B200Array(shape={self.shape}, dtype={self.dtype}, device={self.device})"


class Tensor:
    """`Tensor{T,N,A}`: an array plus one `Index` per dimension (src/Tensor.jl:11-32).
    `data` is a numpy array (host; shape = Julia `size`) or a `B200Array`."""

    def __init__(self, data, inds=()):
        if not isinstance(data, B200Array) and not getattr(data, "_is_block_array", False):   # blocks.BlockArray: the DArray stand-in
            data = np.asarray(data)
        inds = _as_index_list(inds)
        if len(inds) != data.ndim:  # src/Tensor.jl:16-18
            raise ArgumentError(f"ndims(data) [{data.ndim}] must be equal to length(inds) [{len(inds)}]")
        for i in set(inds):          # src/Tensor.jl:20-24
            sizes = {data.shape[d] for d, j in enumerate(inds) if j == i}
            if len(sizes) > 1:
                raise DimensionMismatch("nonuniform size of repeated indices")
        self.data = data
        self._inds = tuple(inds)

    @classmethod
    def _trusted(cls, data, inds_tuple):
        """Result of a backend call: labels and shape are consistent by construction (no constructor checks)."""
        self = object.__new__(cls)
        self.data, self._inds = data, inds_tuple
        return self

    # accessors the backend shim needs (src/Tensor.jl:69,144,156-158,241-242)
    @property
    def inds(self):
        return list(self._inds)

    @property
    def parent(self):
        return self.data

    @property
    def shape(self):
        return tuple(self.data.shape)

    @property
    def ndim(self):
        return self.data.ndim

    @property
    def dtype(self):
        return np.dtype(self.data.dtype)

    def dim(self, ind) -> int:
        ind = ind if isinstance(ind, Index) else Index(ind)
        return self._inds.index(ind)

    def size(self, ind=None):
        if ind is None:
            return self.shape
        return self.shape[self.dim(ind)]

    @property
    def on_device(self) -> bool:
        return isinstance(self.data, B200Array)

    def to_device(self, device=None, non_blocking=False) -> "Tensor":
        if self.on_device:
            return self
        return Tensor(B200Array.from_host(self.data, device, non_blocking=non_blocking), self._inds)

    def to_host(self) -> "Tensor":
        if not self.on_device:
            return self
        return Tensor(self.data.to_host(), self._inds)

    def permutedims(self, perm, flags: int = 0) -> "Tensor":
        """`permutedims(t, perm)` by positions or by `Index` (src/Tensor.jl:302-319). Device tensors go
        through the K1 permute kernel; host tensors are plain container sugar (not on the hot path).
        `flags`: MB200_PERMUTE_TMA (2) / MB200_PERMUTE_NO_TMA (4) pick the transposition kernel (A/B measurements)."""
        if self.ndim == 0:
            return self
        perm = list(perm)
        if perm and not isinstance(perm[0], (int, np.integer)):
            perm = [self._inds.index(i if isinstance(i, Index) else Index(i)) for i in perm]
        new_inds = [self._inds[p] for p in perm]
        if self.on_device:
            src = self.data
            dst = B200Array([src.shape[p] for p in perm], src.dtype, src.device)
            h = _lib.Handle.get(src.device)
            _lib.check(_lib.lib().mb200_permute(h.ptr, C.c_void_p(dst.ptr), C.c_void_p(src.ptr),
                                                _lib.dtype_enum(src.dtype), src.ndim, _lib.i64(src.shape),
                                                _lib.i32(perm), flags))
            return Tensor(dst, new_inds)
        return Tensor(_lib.fortran(np.transpose(self.data, perm)), new_inds)

    def _host_aligned_to(self, other: "Tensor"):
        a = self.to_host().data
        b = other.to_host().data
        if sorted(map(repr, self._inds)) != sorted(map(repr, other._inds)):
            return None, None
        perm = findperm(self._inds, other._inds)
        return a, np.transpose(b, perm)

    def isequal(self, other) -> bool:
        """`isequal(a::Tensor, b::Tensor)` modulo index order (src/Tensor.jl:107-111)."""
        if not isinstance(other, Tensor):
            return False
        a, b = self._host_aligned_to(other)
        return a is not None and a.shape == b.shape and bool(np.array_equal(a, b))

    def isapprox(self, other, rtol=None, atol=0.0) -> bool:
        """`isapprox(a::Tensor, b::Tensor)` (src/Tensor.jl:117-121): relative Frobenius-norm check."""
        if not isinstance(other, Tensor):
            return False
        a, b = self._host_aligned_to(other)
        if a is None or a.shape != b.shape:
            return False
        if rtol is None:
            rtol = float(np.sqrt(np.finfo(np.result_type(a.dtype, b.dtype)).eps))
        na, nb, nd = np.linalg.norm(a.ravel()), np.linalg.norm(b.ravel()), np.linalg.norm((a - b).ravel())
        return bool(nd <= max(atol, rtol * max(na, nb)))

    def __repr__(self):
        return f"Tensor(shape={self.shape}, dtype={self.dtype}, inds={list(self._inds)}, device={self.on_device})"
