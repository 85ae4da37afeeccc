import ctypes as C, sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import torch
import muscle_b200 as mb
from muscle_b200 import B200Array, _lib
import bench_kernels as bk
h = _lib.Handle.get()
for (m, n, k) in ((128, 128, 128), (512, 512, 512), (256, 256, 16384)):
    dt = "complex128"
    a, b = bk.dev_rand((k, m), dt, 1), bk.dev_rand((k, n), dt, 2)
    c = B200Array((m, n), dt)
    e = _lib.dtype_enum(dt)
    for _ in range(3):
        _lib.check(mb.lib().mb200_binary_einsum(h.ptr, C.c_void_p(c.ptr), e, 2, _lib.i32([1, 2]), None,
            C.c_void_p(a.ptr), e, 2, _lib.i32([0, 1]), _lib.i64((k, m)), None,
            C.c_void_p(b.ptr), e, 2, _lib.i32([0, 2]), _lib.i64((k, n)), None))
    torch.cuda.synchronize()
