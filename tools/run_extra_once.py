"""One contraction of each late-round-1 kernel for `ncu --set full`: the persistent gather-GEMM (8192 x 8192 x 256 ComplexF64),
the split tail launch + ordered reduce (2048^3 ComplexF64) and the table-driven gather pack feeding the tcgen05 GEMM
(D = 5 PEPS-like ComplexF32, K = 625: no summed extent tiles groups of 8)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_kernels as bk
from muscle_b200 import Tensor, binary_einsum
for ext, ia, ib, ic, dt in [(dict(i=8192, j=8192, k=256), "ki", "kj", "ij", "complex128"),
                            (dict(i=2048, j=2048, k=2048), "ki", "kj", "ij", "complex128"),
                            (dict(l=125, k=5, b=5, m=125, q=5, r=125, z=5), "lkbmz", "mkqrz", "lbqrz", "complex64")]:
    A = Tensor(bk.dev_rand([ext[c] for c in ia], dt, 1), bk.I(ia)); B = Tensor(bk.dev_rand([ext[c] for c in ib], dt, 2), bk.I(ib))
    c = binary_einsum(A, B, out=bk.I(ic))
    torch.cuda.synchronize()
