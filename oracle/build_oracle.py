"""TEST INFRASTRUCTURE ONLY — builds and loads the plain-C loop-nest oracle (oracle/einsum_ref.c).

The reference (Muscle.jl) is pure Julia: there are no C/C++ sources under /root/reference to
compile, so there is no `oracle/_ref/`; this C file is our own restatement.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "einsum_ref.c")
_SO = os.path.join(_HERE, "liboracle_einsum.so")


def build(force: bool = False) -> str:
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(_SRC):
        subprocess.check_call(["gcc", "-O2", "-std=c11", "-shared", "-fPIC", "-o", _SO, _SRC])
    return _SO


_FN = {np.dtype(np.float32): "oracle_einsum_f32", np.dtype(np.float64): "oracle_einsum_f64",
       np.dtype(np.complex64): "oracle_einsum_c64", np.dtype(np.complex128): "oracle_einsum_c128"}


def einsum_loops(inds_c, a, inds_a, b, inds_b):
    """C[inds_c] = Σ A[inds_a]·B[inds_b] by the explicit loop nest (dense column-major)."""
    lib = ctypes.CDLL(build())
    T = np.result_type(a.dtype, b.dtype)
    a = np.asarray(a, dtype=T)
    b = np.asarray(b, dtype=T)
    a = a if a.ndim == 0 else np.asfortranarray(a)
    b = b if b.ndim == 0 else np.asfortranarray(b)
    labels = []
    for i in list(inds_a) + list(inds_b):
        if i not in labels:
            labels.append(i)
    ext = {}
    for arr, inds in ((a, inds_a), (b, inds_b)):
        for d, i in enumerate(inds):
            ext[i] = arr.shape[d]
    extent = (ctypes.c_int64 * len(labels))(*[ext[i] for i in labels])
    def modes(inds):
        return (ctypes.c_int32 * max(1, len(inds)))(*[labels.index(i) for i in inds])
    c = np.zeros(tuple(ext[i] for i in inds_c), dtype=T, order="F")
    fn = getattr(lib, _FN[np.dtype(T)])
    fn.restype = ctypes.c_int
    rc = fn(ctypes.c_int(len(labels)), extent,
            ctypes.c_int(len(inds_a)), modes(inds_a), ctypes.c_void_p(a.ctypes.data),
            ctypes.c_int(len(inds_b)), modes(inds_b), ctypes.c_void_p(b.ctypes.data),
            ctypes.c_int(len(inds_c)), modes(inds_c), ctypes.c_void_p(c.ctypes.data))
    if rc != 0:
        raise RuntimeError("oracle_einsum failed")
    return c


if __name__ == "__main__":
    print(build(force=True))
