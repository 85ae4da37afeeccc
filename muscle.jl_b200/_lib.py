"""ctypes binding of libmuscle_b200.so (include/muscle_b200.h).

There is no Python/CPU implementation behind this module: if the library is missing or no B200 is
usable, the product path raises — loudly — instead of falling back.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
# MB200_LIB_PATH: A/B benchmarking against another build of the same library (tools/); never a fallback
LIB_PATH = os.environ.get("MB200_LIB_PATH") or os.path.join(_HERE, "libmuscle_b200.so")

MAX_MODES = 32

# dtype enum (mb200_dtype_t)
F32, F64, C64, C128 = 0, 1, 2, 3
_NP2ENUM = {np.dtype(np.float32): F32, np.dtype(np.float64): F64,
            np.dtype(np.complex64): C64, np.dtype(np.complex128): C128}
_ENUM2NP = {v: k for k, v in _NP2ENUM.items()}

# status codes (mb200_status_t)
OK, INVALID_ARGUMENT, NOT_SUPPORTED, DIMENSION_MISMATCH, CUDA_ERROR, OUT_OF_MEMORY, INTERNAL_ERROR = range(7)

# kernel families (mb200_path_t)
PATH_AUTO, PATH_DIRECT, PATH_GETT_F64, PATH_SIMT_F32, PATH_TCGEN05_TF32 = range(5)
PATH_NAMES = {PATH_AUTO: "auto", PATH_DIRECT: "direct", PATH_GETT_F64: "gett_f64_dmma",
              PATH_SIMT_F32: "simt_f32", PATH_TCGEN05_TF32: "tcgen05_split"}

SHARD_NONE, SHARD_FREE, SHARD_BATCH, SHARD_SUM = range(4)

# arithmetic of Float32 / ComplexF32 contractions (mb200_compute_type_t)
COMPUTE_DEFAULT, COMPUTE_FP32, COMPUTE_3XTF32 = range(3)


class ArgumentError(ValueError):
    """Julia `ArgumentError` (src/Operations/binary_einsum.jl:53-55, 82-83)."""


class DimensionMismatch(ValueError):
    """Julia `DimensionMismatch` (src/Tensor.jl:23)."""


class B200Error(RuntimeError):
    """CUDA / allocation / internal failure inside libmuscle_b200."""


class PlanInfo(C.Structure):
    _fields_ = [("M", C.c_int64), ("N", C.c_int64), ("K", C.c_int64), ("L", C.c_int64),
                ("swapped", C.c_int32), ("path", C.c_int32), ("compute_dtype", C.c_int32),
                ("n_left", C.c_int32), ("n_right", C.c_int32), ("n_sum", C.c_int32), ("n_batch", C.c_int32),
                ("left", C.c_int32 * MAX_MODES), ("right", C.c_int32 * MAX_MODES),
                ("sum", C.c_int32 * MAX_MODES), ("batch", C.c_int32 * MAX_MODES),
                ("a_kmajor", C.c_int32), ("b_kmajor", C.c_int32),
                ("flops", C.c_double), ("bytes", C.c_double),
                ("tc_eligible", C.c_int32), ("tc_permute_pack", C.c_int32)]


class ShardInfo(C.Structure):
    _fields_ = [("kind", C.c_int32), ("mode", C.c_int32), ("begin", C.c_int64), ("end", C.c_int64),
                ("needs_allreduce", C.c_int32)]


class Comm(C.Structure):
    """mb200_comm_t: the peer buffers of a fused contraction + all-reduce (include/muscle_b200.h)."""
    _fields_ = [("nranks", C.c_int32), ("rank", C.c_int32), ("epoch", C.c_int32),
                ("ws", C.c_void_p * 8), ("c", C.c_void_p * 8), ("flags", C.c_void_p * 8),
                ("mc_ws", C.c_void_p), ("mc_c", C.c_void_p), ("ws_bytes", C.c_size_t), ("flag_bytes", C.c_size_t)]


DIST_CONTRACT, DIST_REDUCE, DIST_WAIT = 1, 2, 4


class Stats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in (
        "launches_total", "launches_direct", "launches_gett_f64", "launches_simt_f32", "launches_tcgen05",
        "launches_permute", "launches_table", "launches_convert", "launches_reduce", "plans_built", "plans_hit",
        "launches_unary", "launches_hadamard", "graph_launches", "launches_svd", "launches_tcgen05_pair")]

    def as_dict(self):
        return {n: int(getattr(self, n)) for n, _ in self._fields_}


_vp, _i, _i32p, _i64p, _sz = C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int64), C.c_size_t

# every symbol include/muscle_b200.h declares, with its argument types
PROTOTYPES = {
    "mb200_version": ([], C.c_int),
    "mb200_last_error_string": ([], C.c_char_p),
    "mb200_device_count": ([C.POINTER(C.c_int)], C.c_int),
    "mb200_create": ([C.POINTER(_vp), _i], C.c_int),
    "mb200_destroy": ([_vp], C.c_int),
    "mb200_set_stream": ([_vp, _vp], C.c_int),
    "mb200_stream_sync": ([_vp], C.c_int),
    "mb200_set_path": ([_vp, _i], C.c_int),
    "mb200_set_compute_type": ([_vp, _i], C.c_int),
    "mb200_get_compute_type": ([_vp, C.POINTER(_i)], C.c_int),
    "mb200_malloc": ([_vp, C.POINTER(_vp), _sz], C.c_int),
    "mb200_free": ([_vp, _vp], C.c_int),
    "mb200_host_alloc": ([C.POINTER(_vp), _sz], C.c_int),
    "mb200_host_free": ([_vp], C.c_int),
    "mb200_memcpy_h2d": ([_vp, _vp, _vp, _sz], C.c_int),
    "mb200_memcpy_d2h": ([_vp, _vp, _vp, _sz], C.c_int),
    "mb200_memset": ([_vp, _vp, _i, _sz], C.c_int),
    "mb200_binary_einsum": ([_vp,
                             _vp, _i, _i, _i32p, _i64p,
                             _vp, _i, _i, _i32p, _i64p, _i64p,
                             _vp, _i, _i, _i32p, _i64p, _i64p], C.c_int),
    "mb200_binary_einsum_host": ([_vp,
                                  _vp, _i, _i, _i32p,
                                  _vp, _i, _i, _i32p, _i64p,
                                  _vp, _i, _i, _i32p, _i64p], C.c_int),
    "mb200_plan_describe": ([_i, _i, _i32p, _i64p,
                             _i, _i, _i32p, _i64p, _i64p,
                             _i, _i, _i32p, _i64p, _i64p,
                             C.POINTER(PlanInfo)], C.c_int),
    "mb200_permute": ([_vp, _vp, _vp, _i, _i, _i64p, _i32p, C.c_uint32], C.c_int),
    "mb200_unary_einsum": ([_vp,
                            _vp, _i, _i, _i32p, _i64p,
                            _vp, _i, _i, _i32p, _i64p, _i64p], C.c_int),
    "mb200_hadamard": ([_vp, _vp, _i,
                        _vp, _i, _i, _i32p, _i64p,
                        _vp, _i, _i, _i32p, _i64p], C.c_int),
    "mb200_svd_thin": ([_vp, _vp, _vp, _vp, _vp, _i, C.c_int64, C.c_int64, C.c_double, _i], C.c_int),
    "mb200_svd_last_info": ([_vp, C.POINTER(_i), C.POINTER(_i), C.POINTER(_i)], C.c_int),
    "mb200_qr_thin": ([_vp, _vp, _vp, _vp, _i, C.c_int64, C.c_int64], C.c_int),
    "mb200_shard_plan": ([_i, _i32p, _i, _i32p, _i64p, _i, _i32p, _i64p, _i, _i, _i,
                          C.POINTER(ShardInfo)], C.c_int),
    "mb200_ipc_export": ([_vp, _vp, C.c_char_p], C.c_int),
    "mb200_ipc_import": ([_vp, C.c_char_p, C.POINTER(_vp)], C.c_int),
    "mb200_ipc_release": ([_vp, _vp], C.c_int),
    "mb200_binary_einsum_scatter": ([_vp,
                                     _i, _i, _i32p,
                                     _vp, _i, _i, _i32p, _i64p, _i64p,
                                     _vp, _i, _i, _i32p, _i64p, _i64p,
                                     C.POINTER(_vp), _i, _i, _i], C.c_int),
    "mb200_reduce_slots": ([_vp, _vp, _vp, _i, C.c_int64, _i], C.c_int),
    "mb200_allreduce_workspace": ([_vp, _i, _i, _i32p,
                                   _i, _i, _i32p, _i64p,
                                   _i, _i, _i32p, _i64p,
                                   _i, C.POINTER(_sz), C.POINTER(_sz)], C.c_int),
    "mb200_binary_einsum_allreduce": ([_vp, _i, _i, _i32p,
                                       _vp, _i, _i, _i32p, _i64p, _i64p,
                                       _vp, _i, _i, _i32p, _i64p, _i64p,
                                       C.POINTER(Comm), _i], C.c_int),
    "mb200_dist_timeline": ([_vp, C.POINTER(C.c_ulonglong)], C.c_int),
    "mb200_graph_begin": ([_vp], C.c_int),
    "mb200_graph_end": ([_vp, C.POINTER(_vp)], C.c_int),
    "mb200_graph_launch": ([_vp, _vp], C.c_int),
    "mb200_graph_destroy": ([_vp], C.c_int),
    "mb200_signal_peers": ([_vp, C.POINTER(_vp), _i, _i, _i], C.c_int),
    "mb200_reduce_slots_wait": ([_vp, _vp, _vp, _i, C.c_int64, _i, _vp, _i], C.c_int),
    "mb200_get_stats": ([_vp, C.POINTER(Stats)], C.c_int),
    "mb200_reset_stats": ([_vp], C.c_int),
}

_lib = None
_lib_lock = threading.Lock()


def lib() -> C.CDLL:
    """Load libmuscle_b200.so (built in-tree by muscle.jl_b200/build.py). Raises if absent."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.exists(LIB_PATH):
                raise ImportError(
                    f"{LIB_PATH} is missing: build it with `python muscle.jl_b200/build.py` "
                    "(there is no Python or CPU fallback for binary_einsum)")
            L = C.CDLL(LIB_PATH)
            for name, (argtypes, restype) in PROTOTYPES.items():
                if os.environ.get("MB200_LIB_PATH") and not hasattr(L, name):
                    continue           # an older build under A/B test
                fn = getattr(L, name)  # AttributeError if the symbol is not exported
                fn.argtypes = argtypes
                fn.restype = restype
            _lib = L
    return _lib


def last_error() -> str:
    return lib().mb200_last_error_string().decode("utf-8", "replace")


def check(status: int) -> None:
    if status == OK:
        return
    msg = last_error()
    if status in (INVALID_ARGUMENT, NOT_SUPPORTED):
        raise ArgumentError(msg)
    if status == DIMENSION_MISMATCH:
        raise DimensionMismatch(msg)
    raise B200Error(f"libmuscle_b200 status {status}: {msg}")


def fortran(x) -> np.ndarray:
    """Column-major contiguous view/copy (Julia Array layout); 0-dim arrays stay 0-dim."""
    x = np.asarray(x)
    return np.ascontiguousarray(x).reshape(()) if x.ndim == 0 else np.asfortranarray(x)


def dtype_enum(dt) -> int:
    dt = np.dtype(dt)
    if dt not in _NP2ENUM:
        raise ArgumentError(f"eltype {dt} is not supported by BackendB200 "
                            "(Float32, Float64, ComplexF32, ComplexF64)")
    return _NP2ENUM[dt]


def dtype_numpy(e: int) -> np.dtype:
    return _ENUM2NP[e]


def i32(seq):
    seq = list(seq)
    return (C.c_int32 * max(1, len(seq)))(*seq)


def i64(seq):
    seq = list(seq)
    return (C.c_int64 * max(1, len(seq)))(*seq)


class Handle:
    """One mb200 handle per (device, host thread) — the header's contract. The stream follows torch's
    current stream (which is itself per thread) when torch is importable, so torch.cuda.Event timing and
    torch allocations are ordered with our launches (torch is plumbing here: device memory, streams,
    torch.distributed). Two Python threads on different torch streams therefore never re-point each
    other's handle between `set_stream` and a launch (ctypes releases the GIL during foreign calls)."""

    _handles: dict = {}
    _lock = threading.Lock()

    def __init__(self, device: int):
        self.device = device
        self._h = C.c_void_p()
        check(lib().mb200_create(C.byref(self._h), device))
        self._stream = 0

    @classmethod
    def get(cls, device: int | None = None) -> "Handle":
        if device is None:
            device = current_device()
        key = (device, threading.get_ident())
        h = cls._handles.get(key)
        if h is None:
            with cls._lock:
                h = cls._handles.get(key)
                if h is None:
                    h = cls._handles[key] = Handle(device)
        h.sync_stream_with_torch()
        return h

    @property
    def ptr(self):
        return self._h

    def set_stream(self, stream_ptr: int):
        if stream_ptr != self._stream:
            check(lib().mb200_set_stream(self._h, C.c_void_p(stream_ptr)))
            self._stream = stream_ptr

    def sync_stream_with_torch(self):
        raw = _raw_stream_fn()
        if raw is not None:
            sp = raw(self.device)              # torch._C._cuda_getCurrentRawStream: one C call, no Stream object
            if sp != self._stream:
                self.set_stream(sp)

    def synchronize(self):
        check(lib().mb200_stream_sync(self._h))

    def set_path(self, path: int):
        check(lib().mb200_set_path(self._h, path))

    def set_compute_type(self, compute_type: int):
        """COMPUTE_DEFAULT (tensor-core split scheme, <= 1e-5), COMPUTE_FP32 (strict FP32 FMAs), COMPUTE_3XTF32."""
        check(lib().mb200_set_compute_type(self._h, compute_type))

    def compute_type(self) -> int:
        v = C.c_int()
        check(lib().mb200_get_compute_type(self._h, C.byref(v)))
        return int(v.value)

    def stats(self) -> dict:
        s = Stats()
        check(lib().mb200_get_stats(self._h, C.byref(s)))
        return s.as_dict()

    def reset_stats(self):
        check(lib().mb200_reset_stats(self._h))


_RAW_STREAM = [False, None]   # [resolved?, torch._C._cuda_getCurrentRawStream or None]


def _raw_stream_fn():
    """torch's current-stream getter when torch sees a GPU (resolved once), else None."""
    if not _RAW_STREAM[0]:
        fn = None
        try:
            import torch
            if torch.cuda.is_available():
                fn = getattr(torch._C, "_cuda_getCurrentRawStream", None)
                if fn is None:
                    fn = lambda d: int(torch.cuda.current_stream(d).cuda_stream)  # noqa: E731
        except ImportError:
            pass
        _RAW_STREAM[0], _RAW_STREAM[1] = True, fn
    return _RAW_STREAM[1]


def current_device() -> int:
    try:
        import torch
        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except ImportError:
        pass
    return int(os.environ.get("LOCAL_RANK", "0"))


def plan_describe(dtype_c, modes_c, dtype_a, modes_a, ext_a, dtype_b, modes_b, ext_b,
                  strides_c=None, strides_a=None, strides_b=None) -> PlanInfo:
    """Host-only planner introspection (no device needed)."""
    info = PlanInfo()
    check(lib().mb200_plan_describe(
        dtype_c, len(modes_c), i32(modes_c), i64(strides_c) if strides_c is not None else None,
        dtype_a, len(modes_a), i32(modes_a), i64(ext_a), i64(strides_a) if strides_a is not None else None,
        dtype_b, len(modes_b), i32(modes_b), i64(ext_b), i64(strides_b) if strides_b is not None else None,
        C.byref(info)))
    return info


def shard_plan(modes_c, modes_a, ext_a, modes_b, ext_b, nranks, rank, prefer_sum=False) -> ShardInfo:
    info = ShardInfo()
    check(lib().mb200_shard_plan(len(modes_c), i32(modes_c), len(modes_a), i32(modes_a), i64(ext_a),
                                 len(modes_b), i32(modes_b), i64(ext_b), nranks, rank,
                                 1 if prefer_sum else 0, C.byref(info)))
    return info
