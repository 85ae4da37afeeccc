"""n-ary executor numbers: (1) a launch-bound network of small tensors — per-evaluation latency of the step-by-step
Python front-end, of ContractionProgram.run (precompiled, arena) and of the CUDA-graph replay; (2) the BASELINE
config-2 MPS-MPO chain at full size through the executor. Writes gpurun_out/network.json."""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import muscle_b200 as mb  # noqa: E402
from muscle_b200 import B200Array, ContractionProgram, Index, Tensor, binary_einsum  # noqa: E402

I = lambda s: [Index(c) for c in s]


def dev_rand(shape, dtype, seed=0):
    g = torch.Generator(device="cuda:0"); g.manual_seed(seed)
    n = int(np.prod(shape))
    cplx = np.dtype(dtype).kind == "c"
    real = torch.float64 if np.dtype(dtype).itemsize // (2 if cplx else 1) == 8 else torch.float32
    t = torch.rand((2 if cplx else 1) * n, dtype=real, device="cuda:0", generator=g) * 2 - 1
    return B200Array.from_torch(t, shape, dtype)


def wall_us(fn, n):
    for _ in range(20):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / n * 1e6


def main():
    out = {}
    # (1) PEPS-patch style network of 10 small tensors (ComplexF64, bond 4): launch-bound
    labels = ["abc", "cde", "efg", "ghi", "bjk", "dkl", "flm", "hmn", "jo", "lo"]
    ext = {c: 4 for c in "abcdefghijklmno"}
    ts = [Tensor(dev_rand([ext[c] for c in ix], "complex128", k), I(ix)) for k, ix in enumerate(labels)]
    prog = ContractionProgram([t.inds for t in ts], [t.shape for t in ts], [t.dtype for t in ts], out=I("ain"))

    def stepwise():
        live = {i: t for i, t in enumerate(ts)}
        nxt = len(ts)
        for (a, b), st in zip(prog.path, [s for s in prog.steps if s["kind"] == "binary"]):
            ta, tb = live.pop(a), live.pop(b)
            keep = [l for l in dict.fromkeys(ta.inds + tb.inds)
                    if l in prog.out or any(l in t.inds for t in live.values())]
            live[nxt] = binary_einsum(ta, tb, out=keep)
            nxt += 1

    cap = prog.capture(ts)
    n_steps = len(prog.steps)
    r = {"tensors": len(ts), "steps": n_steps,
         "stepwise_binary_einsum_us": wall_us(stepwise, 300),
         "program_run_us": wall_us(lambda: prog.run(ts), 300),
         "graph_replay_us": wall_us(cap.replay, 2000)}
    # device time of one replay
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); e0.record()
    for _ in range(200):
        cap.replay()
    e1.record(); torch.cuda.synchronize()
    r["graph_replay_device_us"] = e0.elapsed_time(e1) / 200 * 1e3
    out["small_network"] = r
    print("NETWORK small (10 tensors, bond 4, c128, %d steps): stepwise %.1f us, program.run %.1f us, graph replay %.1f us (device %.1f us)"
          % (n_steps, r["stepwise_binary_einsum_us"], r["program_run_us"], r["graph_replay_us"], r["graph_replay_device_us"]))

    # (2) config-2 chain at full size through the executor (chi=1024, d=2, w=8), ComplexF64
    e2 = dict(a=1024, b=1024, c=1024, e=1024, w=8, v=8, s=2, t=2)
    labels2 = ["awb", "bsc", "wstv", "ate"]
    ts2 = [Tensor(dev_rand([e2[c] for c in ix], "complex128", 10 + k), I(ix)) for k, ix in enumerate(labels2)]
    for name, path in (("reference label order per step (a's-then-b's), fixed path", None),):
        pass
    progc = ContractionProgram([t.inds for t in ts2], [t.shape for t in ts2], [t.dtype for t in ts2], out=I("evc"),
                               path=[(0, 1), (4, 2), (5, 3)])
    res = progc.run(ts2)
    torch.cuda.synchronize()
    times = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); progc.run(ts2, out=res); e1.record(); torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    ms = float(np.min(times))
    out["config2_chain"] = {"ms_best": ms, "ms_mean": float(np.mean(times)), "flops": progc.flops,
                            "tflops_best": progc.flops / ms / 1e9, "arena_bytes": progc.arena_bytes,
                            "intermediate_orders": {str(k): "".join(str(i.tag) for i in v) for k, v in progc.intermediate_orders.items()}}
    print("NETWORK config-2 chain through the executor: %.3f ms  %.2f TFLOP/s  arena %.0f MB  orders %s"
          % (ms, progc.flops / ms / 1e9, progc.arena_bytes / 1e6, out["config2_chain"]["intermediate_orders"]))
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/network.json", "w"), indent=1)


if __name__ == "__main__":
    main()
