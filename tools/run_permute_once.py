import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muscle_b200 as mb
from muscle_b200 import B200Array, _lib
t = torch.rand(2 * 64**4, dtype=torch.float64, device="cuda:0")
src = B200Array.from_torch(t, (64, 64, 64, 64), "complex128")
dst = B200Array((64, 64, 64, 64), "complex128")
h = _lib.Handle.get()
for _ in range(3):
    _lib.check(mb.lib().mb200_permute(h.ptr, C.c_void_p(dst.ptr), C.c_void_p(src.ptr), _lib.C128, 4, _lib.i64((64,) * 4), _lib.i32((1, 3, 0, 2)), 0))
torch.cuda.synchronize()
