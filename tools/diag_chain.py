"""Diagnostic: why is step 2a slower inside the chain than standalone? Times 2a alone, the chain with fresh
outputs, and the chain with preallocated outputs (binary_einsum_)."""
import os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, ROOT)
import muscle_b200 as mb
from muscle_b200 import B200Array, Index, Tensor, binary_einsum, binary_einsum_
sys.path.insert(0, os.path.join(ROOT))
import bench
I = lambda s: [Index(c) for c in s]
inp = bench.make_inputs(0)
dev = {k: Tensor(v[0], I(v[1])).to_device(0) for k, v in inp.items()}
def ev(): return torch.cuda.Event(enable_timing=True)
def run(label, fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    es = []
    t0 = time.perf_counter()
    for _ in range(n):
        e = [ev() for _ in range(4)]
        fn(e); es.append(e)
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    a = np.mean([e[0].elapsed_time(e[1]) for e in es]); b = np.mean([e[1].elapsed_time(e[2]) for e in es]); c = np.mean([e[2].elapsed_time(e[3]) for e in es])
    print(f"{label:<40s} 2a={a:.3f} 2b={b:.3f} 2c={c:.3f} total={a+b+c:.3f} ms   host-side enqueue {host:.3f} ms/step")
def chain_fresh(e=None):
    if e: e[0].record()
    x = binary_einsum(dev["E"], dev["A"], out=I("awsc"))
    if e: e[1].record()
    y = binary_einsum(x, dev["W"], out=I("atvc"))
    if e: e[2].record()
    z = binary_einsum(y, dev["Ab"], out=I("evc"))
    if e: e[3].record()
X = Tensor(B200Array((1024, 8, 2, 1024), "complex128"), I("awsc")); Y = Tensor(B200Array((1024, 2, 8, 1024), "complex128"), I("atvc")); Z = Tensor(B200Array((1024, 8, 1024), "complex128"), I("evc"))
def chain_prealloc(e=None):
    if e: e[0].record()
    binary_einsum_(X, dev["E"], dev["A"])
    if e: e[1].record()
    binary_einsum_(Y, X, dev["W"])
    if e: e[2].record()
    binary_einsum_(Z, Y, dev["Ab"])
    if e: e[3].record()
def only_2a(e=None):
    if e: e[0].record()
    binary_einsum_(X, dev["E"], dev["A"])
    if e: e[1].record(); e[2].record(); e[3].record()
def only_2c(e=None):
    if e: e[0].record(); e[1].record(); e[2].record()
    binary_einsum_(Z, Y, dev["Ab"])
    if e: e[3].record()
run("chain, fresh outputs", chain_fresh)
run("chain, preallocated outputs", chain_prealloc)
run("2a only (prealloc)", only_2a)
run("2c only (prealloc)", only_2c)
run("chain, fresh outputs (again)", chain_fresh, n=30)
