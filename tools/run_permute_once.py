"""One launch of each K1 permute kernel family for `ncu --set full` (register-tile transposition for 16/8/4-byte
elements, the smem-tile persistent kernel through the split writer is covered by run_cfg3_once.py)."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muscle_b200 as mb
from muscle_b200 import B200Array, _lib
h = _lib.Handle.get()
for shape, perm, dt in (((64, 64, 64, 64), (1, 3, 0, 2), "complex128"), ((8192, 4096), (1, 0), "complex64"),
                        ((8192, 8192), (1, 0), "float32"), ((2, 1024, 1024, 8), (1, 0, 3, 2), "complex128")):
    n = int(np.prod(shape))
    real = torch.float64 if dt == "complex128" else torch.float32
    t = torch.rand((1 if dt == "float32" else 2) * n, dtype=real, device="cuda:0")
    src = B200Array.from_torch(t, shape, dt)
    dst = B200Array([shape[p] for p in perm], dt)
    _lib.check(mb.lib().mb200_permute(h.ptr, C.c_void_p(dst.ptr), C.c_void_p(src.ptr), _lib.dtype_enum(dt), len(shape),
                                      _lib.i64(shape), _lib.i32(perm), 0))
torch.cuda.synchronize()
