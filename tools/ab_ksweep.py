"""K sweep at a fixed 4096 x 4096 ComplexF64 output (2048 tiles): plain vs persistent gather-GEMM (MB200_PERSIST=0 / 2)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import bench_kernels as bk
for K in (128, 256, 512, 1024, 2048, 4096):
    r = bk.einsum_case("c128 4096x4096xK=%d" % K, dict(i=4096, j=4096, k=K), "ki", "kj", "ij", "complex128", iters=8)
    print("EINSUM %-30s %8.2f TF/s best %8.2f mean %.3f ms" % (r["name"], r["tflops_best"], r["tflops_mean"], r["ms_mean"]))
for name, ext, ia, ib, ic, dt in [("c128 step 2a", dict(a=1024, w=8, b=1024, s=2, c=1024), "awb", "bsc", "awsc", "complex128"),
                                  ("c128 step 2c", dict(a=1024, t=2, v=8, c=1024, e=1024), "atvc", "ate", "evc", "complex128")]:
    r = bk.einsum_case(name, ext, ia, ib, ic, dt, iters=8)
    print("EINSUM %-30s %8.2f TF/s best %8.2f mean %.3f ms" % (name, r["tflops_best"], r["tflops_mean"], r["ms_mean"]))
