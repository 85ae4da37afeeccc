"""Config-5 per-rank work at 8 ranks (K sliced to 512), on ONE GPU with local buffers standing in for the peers:
separates the kernel-side cost of the fused scatter epilogue + slot reduce from NVLink / barrier effects."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muscle_b200 as mb
from muscle_b200 import B200Array, Index, Tensor, _lib, binary_einsum
I = lambda s: [Index(c) for c in s]
def dev_rand(shape, seed):
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    t = torch.rand(2 * int(np.prod(shape)), dtype=torch.float32, device="cuda:0", generator=g) * 2 - 1
    return B200Array.from_torch(t, shape, "complex64")
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
n, nr = 8, 8
A = Tensor(dev_rand([n] * 7 + [1], 1), I("aebfcgdh")); B = Tensor(dev_rand([1] + [n] * 7, 2), I("hpgqfres"))
ic = "srqpdcba"
h = _lib.Handle.get(); L = mb.lib()
numel = n ** 8; slab = numel // nr
staging = [B200Array((numel,), "complex64") for _ in range(nr)]
arr = (C.c_void_p * nr)(*[s.ptr for s in staging])
ma, mb_, mc = mb.flatten_labels(A.inds, B.inds, I(ic))
out = B200Array((slab,), "complex64")
cfull = B200Array((n,) * 8, "complex64")
def plain():
    _lib.check(L.mb200_binary_einsum(h.ptr, C.c_void_p(cfull.ptr), _lib.C64, 8, _lib.i32(mc), None,
        C.c_void_p(A.data.ptr), _lib.C64, 8, _lib.i32(ma), _lib.i64(A.shape), None,
        C.c_void_p(B.data.ptr), _lib.C64, 8, _lib.i32(mb_), _lib.i64(B.shape), None))
def scatter():
    _lib.check(L.mb200_binary_einsum_scatter(h.ptr, _lib.C64, 8, _lib.i32(mc),
        C.c_void_p(A.data.ptr), _lib.C64, 8, _lib.i32(ma), _lib.i64(A.shape), None,
        C.c_void_p(B.data.ptr), _lib.C64, 8, _lib.i32(mb_), _lib.i64(B.shape), None, arr, nr, 3, slab.bit_length() - 1))
def reduce():
    _lib.check(L.mb200_reduce_slots(h.ptr, C.c_void_p(out.ptr), C.c_void_p(staging[0].ptr), _lib.C64, slab, nr))
print("per-rank slice of config 5 at 8 ranks (M=N=4096, K=512):")
print(f"  contraction, plain store to C : {timeit(plain):.4f} ms")
print(f"  contraction, scatter epilogue : {timeit(scatter):.4f} ms  (peers = local buffers)")
print(f"  slot reduce (8 x 16.8 MB)     : {timeit(reduce):.4f} ms")
h.reset_stats(); plain(); print(h.stats())
