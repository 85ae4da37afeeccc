"""Per-sweep cost and sweeps-to-convergence of the Jacobi SVD kernels (cluster vs grid: MB200_SVD_CLUSTER)."""
import sys
sys.path.insert(0, "."); sys.path.insert(0, "tools")
import numpy as np, torch
import bench_kernels as bk
from muscle_b200 import Index, Tensor, tensor_svd_thin
I = lambda s: [Index(c) for c in s]
for n, dt in ((64, "complex128"), (128, "complex128"), (128, "complex64"), (256, "complex64")):
    A = Tensor(bk.dev_rand((n, n), dt), I("ab"))
    ref = np.linalg.svd(A.to_host().data, compute_uv=False)
    row = []
    for ms in (1, 2, 4, 6, 8, 10, 12, 30):
        fn = lambda: tensor_svd_thin(A, inds_u=I("a"), ind_s=Index("s"), max_sweeps=ms)
        best, _ = bk.timeit(fn, iters=3, warm=1)
        s = fn()[1].to_host().data
        err = np.linalg.norm(np.sort(s)[::-1] - ref) / np.linalg.norm(ref)
        row.append("%d:%.2fms/%.0e" % (ms, best, err))
    print("SVDPROBE %dx%d %s  " % (n, n, dt) + "  ".join(row))
