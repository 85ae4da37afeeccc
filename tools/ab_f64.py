"""A/B helper: Float64 gather-GEMM variants (MB200_F64_PAIRS=0: CoreD, 1: k-pair core 128x128, 2: k-pair core 128x64)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import bench_kernels as bk
for name, ext, ia, ib, ic, dt in [("f64 8192^3 k-major", dict(i=8192, j=8192, k=8192), "ki", "kj", "ij", "float64"),
                                  ("f64 4096x4096x1024 k-major", dict(i=4096, j=4096, k=1024), "ki", "kj", "ij", "float64"),
                                  ("f64 rank4 k-major", dict(i=64, j=64, k=64, l=64, m=64, n=64), "klij", "klmn", "ijmn", "float64"),
                                  ("f64 2048^3 k-major", dict(i=2048, j=2048, k=2048), "ki", "kj", "ij", "float64"),
                                  ("f64 8192^3 m-major (CoreD)", dict(i=8192, j=8192, k=8192), "ik", "jk", "ij", "float64")]:
    r = bk.einsum_case(name, ext, ia, ib, ic, dt, iters=5)
    print("EINSUM %-30s %8.2f TF/s best %8.2f mean %.3f ms" % (name, r["tflops_best"], r["tflops_mean"], r["ms_mean"]))
