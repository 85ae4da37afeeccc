"""One launch of the unary_einsum / hadamard / Jacobi-SVD kernels for `ncu --set full`."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import muscle_b200 as mb
from muscle_b200 import Index, Tensor
from bench_kernels import dev_rand
I = lambda s: [Index(c) for c in s]
a = Tensor(dev_rand((1024, 16, 1024), "complex128", 1), I("lpr"))
s = Tensor(dev_rand((1024,), "complex128", 2), I("r"))
mb.hadamard(a, s)
mb.hadamard(a, a)
x = Tensor(dev_rand((256, 64, 256, 64), "complex128", 3), I("abcd"))
mb.unary_einsum(x, out=I("ac"))
mb.unary_einsum(Tensor(dev_rand((16384, 1024), "complex128", 4), I("ab")), out=I("b"))
mb.unary_einsum(Tensor(dev_rand((1024, 16384), "complex128", 5), I("ab")), out=[])
torch.cuda.synchronize()
