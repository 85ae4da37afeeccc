"""One contraction of each tcgen05 flavour for `ncu --set full`: config 3 (ComplexF32, batched), config 5
(ComplexF32 rank-8) and an 8192^3 Float32 GEMM (128 x 256 tiles)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muscle_b200 import B200Array, Index, Tensor, binary_einsum
I = lambda s: [Index(c) for c in s]
def dev_rand(shape, seed, dt="complex64"):
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    t = torch.rand((2 if dt == "complex64" else 1) * int(np.prod(shape)), dtype=torch.float32, device="cuda:0", generator=g) * 2 - 1
    return B200Array.from_torch(t, shape, dt)
A = Tensor(dev_rand((256, 8, 8, 256, 8), 1), I("lkbmz")); B = Tensor(dev_rand((256, 8, 8, 256, 8), 2), I("mkqrz"))
c = binary_einsum(A, B, out=I("lbqrz"))
torch.cuda.synchronize()
A5 = Tensor(dev_rand((8,) * 8, 3), I("aebfcgdh")); B5 = Tensor(dev_rand((8,) * 8, 4), I("hpgqfres"))
c = binary_einsum(A5, B5, out=I("srqpdcba"))
torch.cuda.synchronize()
Af = Tensor(dev_rand((8192, 8192), 5, "float32"), I("ki")); Bf = Tensor(dev_rand((8192, 8192), 6, "float32"), I("kj"))
c = binary_einsum(Af, Bf, out=I("ij"))
torch.cuda.synchronize()
