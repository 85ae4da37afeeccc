"""Sustained loop of one contraction with nvidia-smi sampled alongside: SM clock, power draw and throttle reasons while the
tcgen05 kernels run (is the kernel power-bound?). Usage: python tools/power_probe.py [cfg3|f32|c64]"""
import os, subprocess, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import torch
import bench_kernels as bk
from muscle_b200 import Tensor, binary_einsum
which = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
ext, ia, ib, ic, dt = {"cfg3": (dict(l=256, k=8, b=8, m=256, q=8, r=256, z=8), "lkbmz", "mkqrz", "lbqrz", "complex64"),
                       "c64": (dict(i=8192, j=8192, k=8192), "ki", "kj", "ij", "complex64"),
                       "f32": (dict(i=8192, j=8192, k=8192), "ki", "kj", "ij", "float32"),
                       "c128": (dict(i=4096, j=4096, k=4096), "ki", "kj", "ij", "complex128")}[which]
A = Tensor(bk.dev_rand([ext[c] for c in ia], dt, 1), bk.I(ia)); B = Tensor(bk.dev_rand([ext[c] for c in ib], dt, 2), bk.I(ib))
samples, stop = [], False
def sampler():
    while not stop:
        r = subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,clocks.max.sm,power.draw,power.limit,temperature.gpu,clocks_event_reasons.active",
                            "--format=csv,noheader,nounits", "-i", "0"], capture_output=True, text=True)
        samples.append(r.stdout.strip())
        time.sleep(0.05)
th = threading.Thread(target=sampler); th.start()
time.sleep(0.5)
for _ in range(3): binary_einsum(A, B, out=bk.I(ic))
torch.cuda.synchronize()
t0 = time.time(); n = 0
while time.time() - t0 < 4.0:
    for _ in range(20): binary_einsum(A, B, out=bk.I(ic))
    torch.cuda.synchronize(); n += 20
dt_s = time.time() - t0
stop = True; th.join()
import numpy as np
flops = (8.0 if "complex" in dt else 2.0) * float(np.prod([float(v) for v in ext.values()]))
print("%s: %d calls in %.2f s -> %.1f TFLOP/s sustained" % (which, n, dt_s, flops * n / dt_s / 1e12))
for s in samples[::max(1, len(samples) // 12)]: print("  sm_mhz,max,power_w,limit_w,temp,reasons:", s)
