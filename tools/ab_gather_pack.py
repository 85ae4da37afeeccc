"""Shapes whose summed extents cannot tile groups of 8 k (odd bond dimensions, K = 100): tcgen05 through the table-driven
gather pack (auto) against the FFMA gather-GEMM they used before (path forced)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import bench_kernels as bk
import muscle_b200 as mb
from muscle_b200 import _lib
h = _lib.Handle.get()
CASES = [("c64 D=5 PEPS-like", dict(l=125, k=5, b=5, m=125, q=5, r=125, z=5), "lkbmz", "mkqrz", "lbqrz", "complex64"),
         ("c64 D=7 PEPS-like", dict(l=343, k=7, b=7, m=343, q=7, r=343, z=7), "lkbmz", "mkqrz", "lbqrz", "complex64"),
         ("c64 chi=1000 d=3 K=3000", dict(a=1000, s=3, b=1000, c=1000), "bsa", "bsc", "ac", "complex64"),
         ("c64 K=100 chi=2048", dict(i=2048, j=2048, k=100), "ki", "kj", "ij", "complex64"),
         ("f32 4096^2 K=1001", dict(i=4096, j=4096, k=1001), "ki", "kj", "ij", "float32")]
for label, path in (("tcgen05+gather", mb.PATH_AUTO), ("FFMA", mb.PATH_SIMT_F32)):
    h.set_path(path)
    for name, ext, ia, ib, ic, dt in CASES:
        h.reset_stats()
        r = bk.einsum_case(name, ext, ia, ib, ic, dt, iters=6)
        print("EINSUM %-14s %-26s %8.2f TF/s best %8.2f mean %.3f ms  tcgen05=%d" % (label, name, r["tflops_best"], r["tflops_mean"], r["ms_mean"], h.stats()["launches_tcgen05"] > 0))
h.set_path(mb.PATH_AUTO)
