#!/usr/bin/env python
"""bench.py — `binary_einsum` effective TFLOP/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: Muscle's BackendBase restated (oracle)

Workload (configs[1] of BASELINE.json): MPS–MPO transfer contraction, ComplexF64, bond χ=1024, physical
d=2, MPO bond w=8 — one "step" is the whole three-contraction chain
    T  = E[a,w,b]·A[b,s,c]      → [a,w,s,c]   (8192×2048×1024 GEMM-equivalent)
    T' = T·W[w,s,t,v]           → [a,t,v,c]   (1048576×16×16, HBM-bound, a real output permutation)
    E' = T'·conj-site Ā[a,t,e]  → [e,v,c]     (8192×1024×2048)
with synthetic random tensors (NumPy default_rng, seeds 2000..2003). At N GPUs every rank contracts its own
independent chain (weak scaling, no data-path collective); the north star's sharded configs (config 4
free-index shard, config 5 summed-index slice + NCCL all_reduce) are measured after the main timed region
and reported under "sharded_configs" in the same JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

CHI, D, W = 1024, 2, 8
# nominal FP64 tensor peak: 148 SMs x 64 FMA/clk/SM x 2 flop x 1.965 GHz. MEASURED_PEAKS.json carries no FP64
# figure and B200_PROFILING.md states no FP64 fallback; cuBLAS ZGEMM measured on this pool reaches 36.8 TF/s
# (profiles/peaks_r01.json), so the nominal number is a tight ceiling.
FP64_TENSOR_PEAK_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12
TF32_DENSE_PEAK_TFLOPS = 148 * 2048 * 2 * 1.965e9 / 1e12   # 128x256x8 tf32 MMA per 128 clk per SM -> 1191 TFLOP/s nominal
FP64_PEAK_SOURCE = ("nominal FP64 DMMA peak 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s (MEASURED_PEAKS.json has "
                    "no FP64 entry, B200_PROFILING.md no FP64 fallback); cuBLAS ZGEMM 4096^3 measured on this pool: "
                    "36.8 TFLOP/s (profiles/peaks_r01.json)")


def chain_specs(chi=CHI, d=D, w=W):
    """(name, inds_a, inds_b, inds_c) of the three contractions and the tensor shapes."""
    shapes = {"E": ("awb", (chi, w, chi)), "A": ("bsc", (chi, d, chi)), "W": ("wstv", (w, d, d, w)),
              "Ab": ("ate", (chi, d, chi))}
    steps = [("2a", "awb", "bsc", "awsc"), ("2b", "awsc", "wstv", "atvc"), ("2c", "atvc", "ate", "evc")]
    ext = dict(a=chi, b=chi, c=chi, e=chi, w=w, v=w, s=d, t=d)
    flops = []
    for _, ia, ib, ic in steps:
        labels = set(ia) | set(ib)
        flops.append(8.0 * float(np.prod([ext[c] for c in labels], dtype=np.float64)))
    return shapes, steps, flops


def make_inputs(rank=0, chi=CHI, d=D, w=W):
    shapes, _, _ = chain_specs(chi, d, w)
    out = {}
    for k, (name, (inds, shape)) in enumerate(shapes.items()):
        rng = np.random.default_rng(2000 + k + 10 * rank)
        x = rng.uniform(-1, 1, size=shape) + 1j * rng.uniform(-1, 1, size=shape)
        out[name] = (np.asfortranarray(x.astype(np.complex128)), inds)
    return out


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def mark(self):
        """Samples before this point (warm-up) are ignored."""
        try:
            self.skip = sum(1 for _ in open(self.path))
        except Exception:
            self.skip = 0

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for ln, line in enumerate(open(self.path)):
            if ln < getattr(self, "skip", 0):
                continue
            p = [x.strip() for x in line.split(",")]
            if len(p) < 9:
                continue
            try:
                sm.append(float(p[1])); mx.append(float(p[2])); pw.append(float(p[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------
def cpu_chain(inputs, steps):
    """Muscle's default host path restated (oracle): BackendBase TTGT = permutedims copies + BLAS gemm +
    permutedims (src/Operations/binary_einsum.jl:76-96), all host threads."""
    from oracle import binary_einsum_base
    t = {"awb": inputs["E"][0], "bsc": inputs["A"][0], "wstv": inputs["W"][0], "ate": inputs["Ab"][0]}
    cur = None
    for _, ia, ib, ic in steps:
        a = cur if cur is not None else t[ia]
        cur = binary_einsum_base(list(ic), a, list(ia), t[ib], list(ib))
    return cur


def blas_threads():
    try:
        from threadpoolctl import threadpool_info
        n = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(n) if n else os.cpu_count()
    except Exception:
        return os.cpu_count()


def time_cpu(steps_k, warmup, budget_s=150.0):
    """Bounded sample of the same workload: shrink χ if the full chain would not finish in the budget."""
    chi = CHI
    while True:
        inputs = make_inputs(0, chi)
        _, steps, flops = chain_specs(chi)
        t0 = time.perf_counter()
        cpu_chain(inputs, steps)
        first = time.perf_counter() - t0
        if first * (steps_k + warmup) <= budget_s or chi <= 128:
            break
        chi //= 2
    for _ in range(max(0, warmup - 1)):
        cpu_chain(inputs, steps)
    times = []
    for _ in range(steps_k):
        t0 = time.perf_counter()
        cpu_chain(inputs, steps)
        times.append(time.perf_counter() - t0)
    mean = float(np.mean(times))
    sample = (f"full chain chi={chi} d={D} w={W} (3 contractions, {sum(flops) / 1e9:.1f} GFLOP), {steps_k} passes"
              if chi == CHI else
              f"same chain at reduced bond chi={chi} ({sum(flops) / 1e9:.1f} GFLOP per pass; full size did not fit the time budget), {steps_k} passes")
    return sum(flops) / mean / 1e12, mean, sample


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tf, mean, sample = time_cpu(args.steps, args.warmup)
    cores = blas_threads()
    line = {
        "impl": "reference", "metric": "binary_einsum effective TFLOP/s", "value": tf, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": f"MPS-MPO transfer contraction ComplexF64 chi={CHI} d={D} w={W} (configs[1]), 3-step chain",
                   "note": "CPU restatement of Muscle BackendBase (permutedims + OpenBLAS zgemm + permutedims); Julia unavailable"},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import torch.distributed as dist
    import muscle_b200 as mb
    from muscle_b200 import B200Array, Index, Tensor, binary_einsum

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    I = lambda s: [Index(c) for c in s]
    h = mb.Handle.get(local)

    shapes, steps, flops = chain_specs()
    step_flops = float(sum(flops))
    inputs = make_inputs(rank)
    # pinned host copies (e2e source) and a pinned result buffer
    pinned = {}
    for name, (arr, inds) in inputs.items():
        t = torch.empty(arr.size * 2, dtype=torch.float64).pin_memory()
        view = t.numpy().view(np.complex128).reshape(arr.shape, order="F")
        view[...] = arr
        pinned[name] = (view, inds, t)
    res_pin = torch.empty(CHI * W * CHI * 2, dtype=torch.float64).pin_memory()
    res_view = res_pin.numpy().view(np.complex128).reshape((CHI, W, CHI), order="F")

    dev = {name: Tensor(arr, I(inds)).to_device(local) for name, (arr, inds) in inputs.items()}

    def chain(t):
        x = binary_einsum(t["E"], t["A"], out=I("awsc"))
        y = binary_einsum(x, t["W"], out=I("atvc"))
        return binary_einsum(y, t["Ab"], out=I("evc"))

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # the sampler's NVML start-up must not land inside the timed region (it stalls launches for a few ms)
    sampler = ClockSampler(local)
    sampler.start()
    for _ in range(max(args.warmup, 3)):
        chain(dev)
    barrier()

    # ---- timed region: device-resident inputs ------------------------------------------------------
    K = args.steps
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(K)]
    sampler.mark()
    h.reset_stats()
    barrier()
    t_wall0 = time.perf_counter()
    for k in range(K):
        ev[k][0].record()
        x = binary_einsum(dev["E"], dev["A"], out=I("awsc"))
        ev[k][1].record()
        y = binary_einsum(x, dev["W"], out=I("atvc"))
        ev[k][2].record()
        z = binary_einsum(y, dev["Ab"], out=I("evc"))
        ev[k][3].record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t_wall0
    stats = h.stats()
    total_ms = ev[0][0].elapsed_time(ev[K - 1][3])
    clocks = sampler.stop()
    barrier()
    if world > 1:
        tmax = torch.tensor([total_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        total_ms = float(tmax.item())
    ms_per_step = total_ms / K
    value = world * step_flops / (ms_per_step * 1e-3) / 1e12

    t2a = float(np.mean([ev[k][0].elapsed_time(ev[k][1]) for k in range(K)]))
    t2b = float(np.mean([ev[k][1].elapsed_time(ev[k][2]) for k in range(K)]))
    t2c = float(np.mean([ev[k][2].elapsed_time(ev[k][3]) for k in range(K)]))
    # dominant kernel: gett_kernel<CoreZ<128,64,...>> (steps 2a and 2c, one launch each, same flops)
    dom_ms = 0.5 * (t2a + t2c)
    dom_flops = 0.5 * (flops[0] + flops[2])
    achieved = dom_flops / (dom_ms * 1e-3) / 1e12
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(tpath):
        try:
            traffic = json.load(open(tpath)).get("gett_z_128x64_dram_bytes_per_launch")
        except Exception:
            traffic = None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    bytes_2b = 16.0 * (CHI * W * D * CHI + W * D * D * W + CHI * D * W * CHI)   # read T + W, write T'
    roofline = {"bound": "tensor", "kernel": "gett_kernel<CoreZ<128,64,32,32,32,2>>: ComplexF64 4M on DMMA.8x8x4 (steps 2a, 2c)",
                "achieved": achieved, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s", "frac": achieved / FP64_TENSOR_PEAK_TFLOPS,
                "traffic": traffic, "peak_source": FP64_PEAK_SOURCE,
                "flops_per_launch": dom_flops, "ms_per_launch": dom_ms,
                "share_of_step": (t2a + t2c) / (t2a + t2b + t2c)}
    roofline_2b = {"bound": "hbm", "kernel": "stream_kernel<CoreZ<64,16,16,16,8,2>> (step 2b, N=K=16: persistent, B resident in smem, tile pipeline)",
                   "achieved": bytes_2b / (t2b * 1e-3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
                   "frac": bytes_2b / (t2b * 1e-3) / 1e9 / hbm_peak, "bytes_per_launch": bytes_2b, "ms_per_launch": t2b,
                   "peak_source": "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in peaks else "fallback 6650 GB/s"}

    # ---- e2e: public API with HOST (pinned) buffers; every step uploads its inputs and downloads its result -----
    # Steps are pipelined over three streams (upload / contract / download) so the PCIe copies of neighbouring
    # steps overlap the contraction; all copies of all timed steps are inside the timed region.
    Ke = max(3, min(K, 50))   # as many pipelined steps as the device-resident measurement: the un-overlapped first upload and
                              # last download (6 ms together) are inside the timed region and amortise over Ke steps
    h2d = sum(v[0].nbytes for v in pinned.values())
    d2h = res_view.nbytes
    res_pins = [torch.empty(CHI * W * CHI * 2, dtype=torch.float64).pin_memory() for _ in range(2)]
    res_views = [t.numpy().view(np.complex128).reshape((CHI, W, CHI), order="F") for t in res_pins]
    s_in, s_comp, s_out = torch.cuda.Stream(local), torch.cuda.Stream(local), torch.cuda.Stream(local)

    def e2e_run(nsteps, depth=3):
        keep = []                                    # sliding window: a step's device buffers live until its
        for k in range(nsteps):                      # download has finished, then go back to the caching allocator
            if len(keep) >= depth:
                keep[0][2].synchronize()
                keep.pop(0)
            with torch.cuda.stream(s_in):
                t = {name: Tensor(view, I(inds)).to_device(local, non_blocking=True) for name, (view, inds, _) in pinned.items()}
                ev_in = torch.cuda.Event(); ev_in.record(s_in)
            with torch.cuda.stream(s_comp):
                s_comp.wait_event(ev_in)
                z = chain(t)
                ev_c = torch.cuda.Event(); ev_c.record(s_comp)
            with torch.cuda.stream(s_out):
                s_out.wait_event(ev_c)
                z.data.to_host(out=res_views[k & 1], non_blocking=True)
                ev_o = torch.cuda.Event(); ev_o.record(s_out)
            keep.append((t, z, ev_o))
        for st in (s_in, s_comp, s_out):
            st.synchronize()

    def e2e_serial_step():
        t = {name: Tensor(B200Array.from_host(view, local), I(inds)) for name, (view, inds, _) in pinned.items()}
        chain(t).data.to_host(out=res_view)

    e2e_run(6)
    check = float(np.abs(res_views[1] - chain(dev).to_host().data).max())
    e2e_serial_step()
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        e2e_serial_step()
    e1.record()
    torch.cuda.synchronize()
    serial_ms = e0.elapsed_time(e1) / 3
    barrier()
    t0 = time.perf_counter()
    e2e_run(Ke)
    e2e_ms = (time.perf_counter() - t0) * 1e3 / Ke
    if world > 1:
        tmax = torch.tensor([e2e_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
        e2e_ms = float(tmax.item())
    e2e = {"value": world * step_flops / (e2e_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d,
           "d2h_bytes_per_step": d2h, "ms_per_step": e2e_ms, "steps": Ke,
           "serial_ms_per_step": serial_ms, "serial_value": world * step_flops / (serial_ms * 1e-3) / 1e12,
           "max_abs_diff_vs_device_resident_result": check,
           "timing": "host wall clock around Ke pipelined steps, all streams synchronised on both sides",
           "api": "Tensor(pinned numpy).to_device(non_blocking) -> 3x binary_einsum -> .to_host(non_blocking); steps "
                  "pipelined over upload/contract/download streams (serial_* = the same without pipelining)"}

    # ---- the north star's sharded configs ----------------------------------------------------------------
    sharded = None
    if not args.skip_sharded:
        sharded = run_sharded_configs(args, world, rank, local)

    # ---- CPU baseline (rank 0, N=1 only) -------------------------------------------------------------------
    cpu = None
    if world == 1 and not args.skip_cpu:
        tf, mean, sample = time_cpu(3, 1, budget_s=40.0)
        cpu = {"value": tf, "unit": "TFLOP/s", "cores": blas_threads(), "kind": "port", "sample": sample,
               "ms_per_step": mean * 1e3,
               "what": "CPU restatement of Muscle BackendBase (numpy permutedims copies + OpenBLAS zgemm); Julia unavailable"}

    if rank == 0:
        line = {
            "metric": "binary_einsum effective TFLOP/s", "value": value, "unit": "TFLOP/s", "n_gpus": world,
            "steps": K, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"MPS-MPO transfer contraction ComplexF64 chi={CHI} d={D} w={W} (BASELINE.json configs[1]); "
                                   "step = 3-contraction chain 2a,2b,2c; one independent chain per GPU",
                       "arithmetic": "ComplexF64 as 4M real products on FP64 tensor cores (DMMA.8x8x4)",
                       "flops_per_step": step_flops, "l2": "no flush: each step streams 1.4 GB of operands/intermediates (> 126 MB L2)",
                       "parallelism": f"{world} independent replicas, no collective" if world > 1 else "single GPU"},
            "pct_of_fp64_tensor_peak": 100.0 * value / world / FP64_TENSOR_PEAK_TFLOPS,
            "step_breakdown_ms": {"2a": t2a, "2b": t2b, "2c": t2c, "wall_ms_per_step": t_wall / K * 1e3},
            "roofline": roofline, "roofline_2b": roofline_2b, "e2e": e2e, "cpu_baseline": cpu,
            "gpu_launches": int(stats["launches_total"]),
            "gpu_launch_breakdown": {k: v for k, v in stats.items() if v},
            "clocks": clocks, "sharded_configs": sharded,
        }
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def run_sharded_configs(args, world, rank, local):
    """config 4 (free-index shard, no collective, strong scaling) and config 5 (summed-index slice + NCCL
    all_reduce). Synthetic operands are generated on the device (uniform[-1,1)); every rank builds only its
    slab. Times are CUDA-event, max over ranks."""
    import torch
    import torch.distributed as dist
    import muscle_b200 as mb
    from muscle_b200 import B200Array, Index, Tensor, binary_einsum
    from muscle_b200.dist import all_reduce_sum

    I = lambda s: [Index(c) for c in s]
    out = {}

    def dev_rand(shape, dtype, seed):
        g = torch.Generator(device=f"cuda:{local}")
        g.manual_seed(seed)
        n = int(np.prod(shape))
        real = torch.float64 if dtype == "complex128" else torch.float32
        t = torch.rand(2 * n, dtype=real, device=f"cuda:{local}", generator=g) * 2 - 1
        return B200Array.from_torch(t, shape, dtype)

    def timed(fn, iters):
        fn()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / iters
        if world > 1:
            t = torch.tensor([ms], device=f"cuda:{local}", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        return ms

    # config 4b: rank-6 ComplexF64, extents (32,32,16 | 32,32,16) → 16384^3 GEMM-equivalent, sharded over the
    # slowest free index of C (i, extent 16)
    ext = dict(a=32, b=32, c=16, d=32, e=32, f=16, g=32, h=32, i=16)
    if args.cfg4_small:
        ext = {k: 16 for k in ext}
    ia, ib, ic = "adbecf", "fgdhei", "abcghi"
    lo, hi = ext["i"] * rank // world, ext["i"] * (rank + 1) // world
    ext_loc = dict(ext, i=hi - lo)
    A = Tensor(dev_rand([ext[c] for c in ia], "complex128", 4000), I(ia))          # replicated
    B = Tensor(dev_rand([ext_loc[c] for c in ib], "complex128", 4001 + rank), I(ib))  # this rank's slab
    flops4 = 8.0 * float(np.prod([ext[c] for c in ext], dtype=np.float64))
    ms = timed(lambda: binary_einsum(A, B, out=I(ic)), 2 if not args.cfg4_small else 5)
    out["config4_free_index_shard"] = {
        "workload": f"rank-6 ComplexF64, extents {ext}, 3 summed; C sharded over free index i ({world} slabs), no collective",
        "scaling": "strong", "tflops": flops4 / (ms * 1e-3) / 1e12, "ms": ms, "flops": flops4,
        "pct_of_fp64_tensor_peak_per_gpu": 100.0 * flops4 / (ms * 1e-3) / 1e12 / world / FP64_TENSOR_PEAK_TFLOPS}
    del A, B
    torch.cuda.empty_cache()

    # config 5: rank-8 ComplexF32, dim 8, 4 summed; summed index h sliced over the ranks, partials all-reduced
    n = 8
    ia, ib, ic = "aebfcgdh", "hpgqfres", "srqpdcba"
    hl = max(1, n // world) if world <= n else 1
    exta = [n] * 7 + [hl]
    extb = [hl] + [n] * 7
    A5 = Tensor(dev_rand(exta, "complex64", 5000 + rank), I(ia))
    B5 = Tensor(dev_rand(extb, "complex64", 5100 + rank), I(ib))
    flops5 = 8.0 * float(n ** 12) * (hl * min(world, n) / n)

    def step5():
        c = binary_einsum(A5, B5, out=I(ic))
        if world > 1:
            all_reduce_sum(c)
        return c

    ms = timed(step5, 5)
    ms_gemm = timed(lambda: binary_einsum(A5, B5, out=I(ic)), 5)
    h5 = mb.Handle.get(local).stats()
    fused = None
    if world > 1:
        # the same slice with the reduction fused into the GEMM epilogue: peer-memory stores to the owner's
        # staging slot + local slot sum (reduce-scatter semantics: C stays sharded), checked against the NCCL result
        from muscle_b200.dist import sum_slice_reduce_scatter
        try:
            slab_t = sum_slice_reduce_scatter(A5, B5, I(ic))
            full = step5().data.to_host().reshape(-1, order="F")
            mine = slab_t.data.to_host().reshape(-1, order="F")
            sl = full[rank * mine.size:(rank + 1) * mine.size]
            err = float(np.linalg.norm(mine - sl) / max(np.linalg.norm(sl), 1e-30))
            ms_f = timed(lambda: sum_slice_reduce_scatter(A5, B5, I(ic)), 5)
            fused = {"tflops": flops5 / (ms_f * 1e-3) / 1e12, "ms": ms_f, "rel_err_vs_nccl_allreduce": err,
                     "semantics": "reduce-scatter (each rank ends with its 1/N slab of C)",
                     "how": "tcgen05 epilogue stores each element into the owner rank's staging slot over NVLink peer "
                            "mappings (CUDA IPC); no collective: every rank signals an epoch flag into every rank's flag array "
                            "(st.release.sys after the GEMM), the owner's slot-sum kernel waits for all N flags"}
        except Exception as e:  # noqa: BLE001
            fused = {"error": repr(e)[:300]}
    out["config5_summed_slice_allreduce"] = {
        "workload": f"rank-8 ComplexF32 dim 8, 4 summed; summed index h sliced {min(world, n)}x, partial C (134 MB) all_reduce(SUM) over NCCL",
        "scaling": "strong", "tflops": flops5 / (ms * 1e-3) / 1e12, "ms": ms, "ms_contraction_only": ms_gemm,
        "flops": flops5, "allreduce_bytes": 8 * n ** 8 if world > 1 else 0,
        "kernel": "pack x2 (K1 split writer) + tcgen05 TF32+BF16 split GEMM" if h5["launches_tcgen05"] else "FFMA gather-GEMM",
        "fused_reduce_scatter": fused}
    del A5, B5
    torch.cuda.empty_cache()

    # config 3: PEPS double-layer environment contraction ComplexF32, D=8, chi=256, batch hyperindex beta=8; sharded
    # over the batch index (no collective)
    chi, Dd, beta = 256, 8, 8
    bl = max(1, beta // world) if world <= beta else 1
    A3 = Tensor(dev_rand([chi, Dd, Dd, chi, bl], "complex64", 3000 + rank), I("lkbmz"))
    B3 = Tensor(dev_rand([chi, Dd, Dd, chi, bl], "complex64", 3100 + rank), I("mkqrz"))
    flops3 = 8.0 * float(chi * Dd) ** 3 * (bl * min(world, beta))
    ms = timed(lambda: binary_einsum(A3, B3, out=I("lbqrz")), 5)
    tf3 = flops3 / (ms * 1e-3) / 1e12
    out["config3_peps_batched_c64"] = {
        "workload": f"PEPS double-layer ComplexF32 D=8 chi=256 beta=8 (2048^3 x 8 GEMM-equivalent), batch index sharded {min(world, beta)}x, no collective",
        "scaling": "strong", "tflops": tf3, "ms": ms, "flops": flops3,
        "pct_of_split_ceiling_per_gpu": 100.0 * tf3 / world / (TF32_DENSE_PEAK_TFLOPS / 2.0),
        "ceiling_note": "TF32 + BF16 split x 4M = 8 tf32-equivalent MACs per complex MAC (8 flops): pipe ceiling = TF32 dense peak / 2; "
                        f"TF32 dense peak taken as nominal {TF32_DENSE_PEAK_TFLOPS:.0f} TFLOP/s (cuBLAS TF32 SGEMM measured 694). The kernel runs at "
                        "the 1000 W board power cap (tools/power_probe.py: 982 W, sw_power_cap, SM clock 1.68 GHz), which binds before the pipe"}
    return out


_REAL_STDOUT = None


def emit(line: dict) -> None:
    """The ONE JSON line of this run, written to the process's original stdout."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # stdout must carry exactly one JSON line, but libraries write banners there from C (NCCL prints "NCCL version ..."
    # with printf, NCCL_DEBUG_FILE does not catch it): point fd 1 at stderr for the whole run and keep the original
    # stdout for the JSON line only.
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=50)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-sharded", action="store_true")
    ap.add_argument("--cfg4-small", action="store_true", help="config 4 at dim 16 (4096^3) instead of 16384^3")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
