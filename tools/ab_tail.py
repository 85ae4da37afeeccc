"""A/B helper: ComplexF64 / Float64 shapes with a ragged last wave (MB200_SPLITK=2 disables the split tail launch)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import bench_kernels as bk
for name, ext, ia, ib, ic, dt in [("c128 2048^3", dict(i=2048, j=2048, k=2048), "ki", "kj", "ij", "complex128"),
                                  ("c128 1280x1280x4096", dict(i=1280, j=1280, k=4096), "ki", "kj", "ij", "complex128"),
                                  ("c128 1536^3", dict(i=1536, j=1536, k=1536), "ki", "kj", "ij", "complex128"),
                                  ("c128 2560x1024x2048", dict(i=2560, j=1024, k=2048), "ki", "kj", "ij", "complex128"),
                                  ("c128 chi=1024 d=2 step 2a", dict(a=1024, w=8, b=1024, s=2, c=1024), "awb", "bsc", "awsc", "complex128"),
                                  ("f64 3072^3", dict(i=3072, j=3072, k=3072), "ki", "kj", "ij", "float64"),
                                  ("c64 ffma 1000x1000x1000", dict(i=1000, j=1000, k=1000), "ik", "jk", "ij", "complex64")]:
    r = bk.einsum_case(name, ext, ia, ib, ic, dt, iters=8)
    print("EINSUM %-30s %8.2f TF/s best %8.2f mean %.3f ms" % (name, r["tflops_best"], r["tflops_mean"], r["ms_mean"]))
# persistent gather-GEMM (MB200_PERSIST=0 disables): the bench's K = 1024 steps and config 1
for name, ext, ia, ib, ic, dt in [("c128 step 2c", dict(a=1024, t=2, v=8, c=1024, e=1024), "atvc", "ate", "evc", "complex128"),
                                  ("c128 cfg1 scrambled", dict(i=64, j=64, k=64, l=64, m=64, n=64), "kilj", "nlmk", "mjni", "complex128"),
                                  ("c128 4096^3", dict(i=4096, j=4096, k=4096), "ki", "kj", "ij", "complex128"),
                                  ("c128 8192x8192x256", dict(i=8192, j=8192, k=256), "ki", "kj", "ij", "complex128")]:
    r = bk.einsum_case(name, ext, ia, ib, ic, dt, iters=8)
    print("EINSUM %-30s %8.2f TF/s best %8.2f mean %.3f ms" % (name, r["tflops_best"], r["tflops_mean"], r["ms_mean"]))
