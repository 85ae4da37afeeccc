"""Gate application (quantum-circuit simulation shape): a 1- or 2-qubit gate contracted into an n-qubit state."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muscle_b200 as mb
from muscle_b200 import B200Array, Index, Tensor, _lib, binary_einsum
def dev_rand(shape, dtype, seed):
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    real = torch.float64 if dtype == "complex128" else torch.float32
    t = torch.rand(2 * int(np.prod(shape)), dtype=real, device="cuda:0", generator=g) * 2 - 1
    return B200Array.from_torch(t, shape, dtype)
def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n
nq = 26
labels = [Index(f"q{i}") for i in range(nq)]
for dt in ("complex128", "complex64"):
    psi = Tensor(dev_rand((2,) * nq, dt, 1), labels)
    for qs in ([0], [13], [25], [3, 11], [0, 1], [24, 25], [5, 12, 20]):
        k = len(qs)
        outs = [Index(f"o{i}") for i in range(k)]
        gate = Tensor(dev_rand((2,) * (2 * k), dt, 2), outs + [labels[q] for q in qs])
        ic = list(labels)
        for o, q in zip(outs, qs): ic[q] = o
        h = _lib.Handle.get(); h.reset_stats()
        ms = timeit(lambda: binary_einsum(gate, psi, out=ic))
        st = {k2: v for k2, v in h.stats().items() if v and k2.startswith("launches_") and k2 != "launches_total"}
        nbytes = 2 * psi.data.nbytes
        print(f"{dt:10s} {nq}-qubit state, gate on {str(qs):12s}: {ms:7.3f} ms  {nbytes / ms / 1e6:7.0f} GB/s  {list(st)}")
