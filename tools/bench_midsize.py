"""Mid-size contractions (too few tiles to fill 148 SMs): device time per call through the C ABI (50 calls back to
back between two CUDA events, so neither Python nor launch latency is in the number), ComplexF64 / ComplexF32."""
import ctypes as C
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, torch
import muscle_b200 as mb
from muscle_b200 import B200Array, _lib
import bench_kernels as bk


ROWS = []


def case(m, n, k, dt, reps=50):
    a, b = bk.dev_rand((k, m), dt, 1), bk.dev_rand((k, n), dt, 2)
    c = B200Array((m, n), dt)
    h = _lib.Handle.get()
    e = _lib.dtype_enum(dt)
    args = (h.ptr, C.c_void_p(c.ptr), e, 2, _lib.i32([1, 2]), None,
            C.c_void_p(a.ptr), e, 2, _lib.i32([0, 1]), _lib.i64((k, m)), None,
            C.c_void_p(b.ptr), e, 2, _lib.i32([0, 2]), _lib.i64((k, n)), None)
    fn = mb.lib().mb200_binary_einsum
    for _ in range(5):
        _lib.check(fn(*args))
    torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn(*args)
        e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) / reps)
    flops = (8.0 if np.dtype(dt).kind == "c" else 2.0) * m * n * k
    info = mb.plan_describe(e, [1, 2], e, [0, 1], [k, m], e, [0, 2], [k, n])
    ROWS.append({"m": m, "n": n, "k": k, "dtype": dt, "us": best * 1e3, "tflops": flops / best / 1e9, "path": mb.PATH_NAMES[info.path]})
    print("MID %5dx%5dx%6d %-10s %8.2f TF/s  %9.2f us  path=%s" % (m, n, k, dt, flops / best / 1e9, best * 1e3, mb.PATH_NAMES[info.path]))


for dt in ("complex128", "complex64"):
    for n in (64, 128, 256, 384, 512, 768, 1024, 1536, 2048):
        case(n, n, n, dt)
    for (m, n, k) in ((512, 512, 4096), (256, 256, 16384), (4096, 64, 4096), (1024, 16, 1024), (2048, 2, 2048), (64, 64, 65536)):
        case(m, n, k, dt)

import json, os
os.makedirs("gpurun_out", exist_ok=True)
json.dump({"midsize": ROWS, "note": "device time per call, 50 calls back to back through the C ABI; split-K active (MB200_SPLITK=0 disables)"},
          open("gpurun_out/kernels_midsize.json", "w"), indent=1)
