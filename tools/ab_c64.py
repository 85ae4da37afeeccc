"""A/B helper: time the ComplexF32 configs (3, 5) and plain GEMM shapes with whatever library MB200_LIB_PATH / environment
switches (MB200_SPLIT_SCHEME=3xtf32, MB200_CTA_PAIR=0) select."""
import sys
sys.path.insert(0, ".")
sys.path.insert(0, "tools")
import bench_kernels as bk
for name, ext, ia, ib, ic, dt in [("cfg3 PEPS c64", dict(l=256, k=8, b=8, m=256, q=8, r=256, z=8), "lkbmz", "mkqrz", "lbqrz", "complex64"),
                                  ("cfg5 rank8 c64", {c: 8 for c in "abcdefghpqrs"}, "aebfcgdh", "hpgqfres", "srqpdcba", "complex64"),
                                  ("c64 4096^3 aligned", dict(i=4096, j=4096, k=4096), "ki", "kj", "ij", "complex64"),
                                  ("c64 8192^3 aligned", dict(i=8192, j=8192, k=8192), "ki", "kj", "ij", "complex64"),
                                  ("f32 8192^3 aligned", dict(i=8192, j=8192, k=8192), "ki", "kj", "ij", "float32"),
                                  ("cfg1 c128 scrambled", dict(i=64, j=64, k=64, l=64, m=64, n=64), "kilj", "nlmk", "mjni", "complex128")]:
    r = bk.einsum_case(name, ext, ia, ib, ic, dt, iters=8)
    print("EINSUM %-30s %8.2f TF/s best %8.2f mean %.3f ms" % (name, r["tflops_best"], r["tflops_mean"], r["ms_mean"]))
