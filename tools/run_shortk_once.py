"""One short-K ComplexF64 contraction (4096 x 4096 x 128: 2048 tiles of 4 k-blocks) for a source-level ncu capture: where do the
~13 us per tile round outside the k loop go? (MB200_PERSIST=0 keeps the plain kernel.)"""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_kernels as bk
from muscle_b200 import Tensor, binary_einsum
A = Tensor(bk.dev_rand([128, 4096], "complex128", 1), bk.I("ki")); B = Tensor(bk.dev_rand([128, 4096], "complex128", 2), bk.I("kj"))
for _ in range(2):
    c = binary_einsum(A, B, out=bk.I("ij"))
torch.cuda.synchronize()
