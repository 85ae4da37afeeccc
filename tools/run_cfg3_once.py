import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muscle_b200 import B200Array, Index, Tensor, binary_einsum
I = lambda s: [Index(c) for c in s]
def dev_rand(shape, seed):
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    t = torch.rand(2 * int(np.prod(shape)), dtype=torch.float32, device="cuda:0", generator=g) * 2 - 1
    return B200Array.from_torch(t, shape, "complex64")
A = Tensor(dev_rand((256, 8, 8, 256, 8), 1), I("lkbmz")); B = Tensor(dev_rand((256, 8, 8, 256, 8), 2), I("mkqrz"))
for _ in range(3):
    c = binary_einsum(A, B, out=I("lbqrz"))
torch.cuda.synchronize()
A5 = Tensor(dev_rand((8,) * 8, 3), I("aebfcgdh")); B5 = Tensor(dev_rand((8,) * 8, 4), I("hpgqfres"))
for _ in range(3):
    c = binary_einsum(A5, B5, out=I("srqpdcba"))
torch.cuda.synchronize()
