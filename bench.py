#!/usr/bin/env python
"""bench.py — `binary_einsum` effective TFLOP/s on B200 (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # CPU arm: Muscle's BackendBase restated (oracle)

Headline workload (every N): BASELINE.json configs[3], the "~16k x 16k x 16k effective GEMM" reading — ONE pairwise
contraction of two rank-6 ComplexF64 tensors, extents (32,32,16) per label class, 3 summed labels, labels interleaved in
memory:  C[a,b,c,g,h,i] = sum_{d,e,f} A[a,d,b,e,c,f] * B[f,g,d,h,e,i]   (16384^3, 35.2 TFLOP, 12.9 GB of operands).
It is the largest single-GPU configuration and the one BASELINE names for 1/2/4/8 GPUs: at N GPUs the free label `i`
of B and C is cut into N slabs (mb200_shard_plan), A is replicated, every rank contracts its slab — STRONG scaling, no
data-path collective (Dagger's loop over output blocks, ext/MuscleDaggerExt/binary_einsum.jl:88-105).
A "step" = one `binary_einsum` call through the public API on device-resident operands.

`per_config` carries the other BASELINE configs (1: rank-4 dim 64; 2: MPS-MPO chain; 3: PEPS batched ComplexF32;
4a: rank-6 dim 16; 5: rank-8 ComplexF32) at N = 1 — value, roofline, ncu DRAM traffic, clocks, parity — and at N > 1 the
sharded configs (3 over the batch label, 5 sliced over a summed label + all-reduce) with parity of rank 0's result
against the oracle. Parity is always computed OUTSIDE the timed regions.
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

# ---- workloads ----------------------------------------------------------------------------------------------------
CFG4B = dict(ext=dict(a=32, b=32, c=16, d=32, e=32, f=16, g=32, h=32, i=16), ia="adbecf", ib="fgdhei", ic="abcghi",
             shard="i", dtype="complex128")
WORKLOAD = ("large rank-6 ComplexF64 contraction, extents (32,32,16 | 32,32,16 | 32,32,16), 3 summed labels, "
            "16384^3 GEMM-equivalent (BASELINE.json configs[3]); C[abcghi] = A[adbecf] * B[fgdhei]")
METRIC = "binary_einsum effective TFLOP/s"
CHI, D, W = 1024, 2, 8                      # config 2 (MPS-MPO chain)

# nominal FP64 tensor peak: 148 SMs x 64 FMA/clk/SM x 2 flop x 1.965 GHz. MEASURED_PEAKS.json carries no FP64 figure and
# B200_PROFILING.md states no FP64 fallback; the measured cuBLAS Z/DGEMM rates of this pool (profiles/peaks_r01.json,
# tools/measure_peaks.py) are reported next to it.
FP64_TENSOR_PEAK_TFLOPS = 148 * 64 * 2 * 1.965e9 / 1e12
TF32_DENSE_PEAK_TFLOPS = 148 * 2048 * 2 * 1.965e9 / 1e12   # 128x256x8 tf32 MMA per 128 clk per SM -> 1191 TFLOP/s nominal


def load_json(path, default=None):
    try:
        return json.load(open(path))
    except Exception:
        return default


PEAKS = load_json(os.path.join(ROOT, "MEASURED_PEAKS.json"), {}) or {}
CUBLAS = load_json(os.path.join(ROOT, "profiles", "peaks_r01.json"), {}) or {}
TRAFFIC = load_json(os.path.join(ROOT, "profiles", "ncu_traffic.json"), {}) or {}
HBM_PEAK = float(PEAKS.get("hbm_gbs", 6650.0))
HBM_SRC = "MEASURED_PEAKS.json hbm_gbs (measured)" if "hbm_gbs" in PEAKS else "fallback 6650 GB/s (B200_PROFILING.md)"


def fp64_roofline(kernel, flops, ms, traffic_key=None):
    ach = flops / (ms * 1e-3) / 1e12
    r = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": FP64_TENSOR_PEAK_TFLOPS, "unit": "TFLOP/s",
         "frac": ach / FP64_TENSOR_PEAK_TFLOPS, "traffic": TRAFFIC.get(traffic_key) if traffic_key else None,
         "flops_per_launch": flops, "ms_per_launch": ms,
         "peak_source": "nominal FP64 DMMA peak 148 SM x 64 FMA/clk x 2 x 1.965 GHz = 37.2 TFLOP/s (MEASURED_PEAKS.json has no "
                        "FP64 entry, B200_PROFILING.md no FP64 fallback)"}
    z = (CUBLAS.get("zgemm_4096") or {}).get("burst_tflops")
    if z:
        r["peak_measured_cublas_zgemm"] = z
        r["frac_of_cublas_zgemm"] = ach / z
    return r


def tf32_roofline(kernel, flops, ms, traffic_key=None, real=False):
    """ComplexF32 / Float32 on the tcgen05 split scheme: 8 tf32-equivalent MACs per complex MAC (8 flops) -> pipe ceiling in
    effective flops = TF32 dense peak / 2 (complex), TF32 dense peak / 2 (real: 2 MMAs per 2-flop MAC)."""
    ach = flops / (ms * 1e-3) / 1e12
    peak = TF32_DENSE_PEAK_TFLOPS / 2.0
    r = {"bound": "tensor", "kernel": kernel, "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
         "traffic": TRAFFIC.get(traffic_key) if traffic_key else None, "flops_per_launch": flops, "ms_per_launch": ms,
         "peak_source": f"nominal TF32 dense {TF32_DENSE_PEAK_TFLOPS:.0f} TFLOP/s / 2 (TF32 + BF16 split x 4M = 8 tf32-equivalent MACs per "
                        "complex MAC); the kernel runs at the 1000 W board power cap, which binds before the pipe"}
    t = (CUBLAS.get("sgemm_tf32_8192") or {}).get("burst_tflops")
    if t:
        r["peak_measured_cublas_tf32"] = t
        r["frac_of_cublas_tf32_half"] = ach / (t / 2.0)
    b = PEAKS.get("bf16_tflops_sustained")
    if b:
        r["peak_proxy_bf16_sustained_half"] = b / 2.0 / 2.0     # tf32 = bf16 / 2; split scheme / 2
        r["frac_of_bf16_proxy"] = ach / (b / 4.0)
    return r


def hbm_roofline(kernel, nbytes, ms, traffic_key=None):
    ach = nbytes / (ms * 1e-3) / 1e9
    return {"bound": "hbm", "kernel": kernel, "achieved": ach, "peak": HBM_PEAK, "unit": "GB/s", "frac": ach / HBM_PEAK,
            "traffic": TRAFFIC.get(traffic_key) if traffic_key else None, "bytes_per_launch": nbytes, "ms_per_launch": ms,
            "peak_source": HBM_SRC}


# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed regions (B200_PROFILING.md recipe). One process for the whole
    run; `mark()` returns the current sample count so any region can be summarised afterwards (`window(m0, m1)`).
    `start()` blocks until the first sample line has been written: nvidia-smi's NVML start-up stalls kernel launches for a
    few ms and must not land inside warm-up or a timed region (it inflated SCALE_r01's N = 1 point by 15 %)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, device):
        self.device, self.proc, self.path = device, None, None

    def start(self, wait_s=15.0):
        try:
            f = tempfile.NamedTemporaryFile("w", suffix=".csv", delete=False)
            self.path = f.name
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.device), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "50"], stdout=f, stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None
            return
        t0 = time.time()
        while time.time() - t0 < wait_s and self.mark() < 2:
            time.sleep(0.02)

    def mark(self):
        try:
            return sum(1 for _ in open(self.path))
        except Exception:
            return 0

    def window(self, m0, m1=None):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        sm, mx, pw, reasons = [], [], [], set()
        for ln, line in enumerate(open(self.path)):
            if ln < m0 or (m1 is not None and ln >= m1):
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
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples in this window (region shorter than 50 ms)"]}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "power_w_max": float(max(pw)),
                "samples": len(sm), "reasons": sorted(reasons)}

    def stop(self):
        if not self.proc:
            return
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        try:
            os.unlink(self.path)
        except OSError:
            pass


# ---- CPU arm: the oracle port of Muscle's BackendBase ------------------------------------------------------------------
def set_blas_threads():
    """All host threads for OpenBLAS whatever the launcher exported (torchrun sets OMP_NUM_THREADS=1, which made round 1's
    N > 1 reference arm single-threaded)."""
    n = os.cpu_count() or 1
    try:
        from threadpoolctl import threadpool_limits, threadpool_info
        threadpool_limits(limits=n, user_api="blas")
        got = [p["num_threads"] for p in threadpool_info() if p.get("user_api") == "blas"]
        return max(got) if got else n
    except Exception:
        return n


def cfg4b_sample(frac_a, frac_g, seed=4200):
    """Bounded sample of the headline workload for the CPU arm: the free labels a and g restricted to their first
    frac_a / frac_g values (a slab of C and of both operands; the summed range K = 16384 stays whole)."""
    ext = dict(CFG4B["ext"])
    ext["a"] = max(1, int(ext["a"] * frac_a))
    ext["g"] = max(1, int(ext["g"] * frac_g))
    rng = np.random.default_rng(seed)

    def rnd(labels):
        shape = tuple(ext[c] for c in labels)
        x = np.empty(shape, dtype=np.complex128, order="F")
        x.real = rng.uniform(-1, 1, size=shape)
        x.imag = rng.uniform(-1, 1, size=shape)
        return x
    flops = 8.0 * float(np.prod([ext[c] for c in ext], dtype=np.float64))
    return rnd(CFG4B["ia"]), rnd(CFG4B["ib"]), ext, flops


def time_cpu_cfg4b(steps_k, warmup, budget_s):
    """Muscle's default host path restated (oracle.binary_einsum_base: permutedims copies + OpenBLAS zgemm + permutedims,
    src/Operations/binary_einsum.jl:76-96) on a bounded slab of the 16384^3 contraction, all host threads."""
    from oracle import binary_einsum_base
    cores = set_blas_threads()
    ia, ib, ic = list(CFG4B["ia"]), list(CFG4B["ib"]), list(CFG4B["ic"])
    frac = 0.25
    while True:
        A, B, ext, flops = cfg4b_sample(frac, frac)
        t0 = time.perf_counter()
        binary_einsum_base(ic, A, ia, B, ib)
        first = time.perf_counter() - t0
        if first * (steps_k + warmup) <= budget_s or frac <= 1.0 / 16:
            break
        frac /= 2
    for _ in range(max(0, warmup - 1)):
        binary_einsum_base(ic, A, ia, B, ib)
    times = []
    for _ in range(steps_k):
        t0 = time.perf_counter()
        binary_einsum_base(ic, A, ia, B, ib)
        times.append(time.perf_counter() - t0)
    mean = float(np.mean(times))
    M = ext["a"] * ext["b"] * ext["c"]
    N = ext["g"] * ext["h"] * ext["i"]
    K = ext["d"] * ext["e"] * ext["f"]
    sample = (f"slab of the 16384^3 contraction: free labels a and g restricted to {ext['a']} / {ext['g']} of 32 values "
              f"(GEMM-equivalent {M} x {N} x {K}, {flops / 1e12:.2f} TFLOP per pass incl. the three permutedims copies), "
              f"{steps_k} passes; the full contraction is {int(round(1 / frac ** 2))}x this slab (35.2 TFLOP, ~2 min per pass on the CPU)")
    return flops / mean / 1e12, mean, sample, cores


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    tf, mean, sample, cores = time_cpu_cfg4b(args.steps, max(args.warmup, 1), budget_s=170.0)
    line = {
        "impl": "reference", "metric": METRIC, "value": tf, "unit": "TFLOP/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": mean * 1e3,
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": WORKLOAD,
                   "note": "CPU restatement of Muscle BackendBase (permutedims + OpenBLAS zgemm + permutedims); Julia unavailable"},
        "cpu_baseline": {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": tf, "unit": "TFLOP/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)


# ---- device-side helpers ----------------------------------------------------------------------------------------------
class Ctx:
    pass


def make_ctx():
    import torch
    import torch.distributed as dist
    c = Ctx()
    c.torch, c.dist = torch, dist
    c.world = int(os.environ.get("WORLD_SIZE", "1"))
    c.rank = int(os.environ.get("RANK", "0"))
    c.local = int(os.environ.get("LOCAL_RANK", "0"))
    c.dev = f"cuda:{c.local}"
    return c


def dev_rand(c, shape, dtype, seed, out=None):
    """uniform[-1,1) re/im generated on the device by a seeded torch generator (synthetic data; torch is the allocator)."""
    from muscle_b200 import B200Array
    torch = c.torch
    g = torch.Generator(device=c.dev)
    g.manual_seed(seed)
    n = int(np.prod(shape))
    cplx = dtype in ("complex128", "complex64")
    real = torch.float64 if dtype in ("complex128", "float64") else torch.float32
    m = (2 if cplx else 1) * n
    if out is None:
        out = torch.empty(m, dtype=real, device=c.dev)
    out.uniform_(-1, 1, generator=g)
    return B200Array.from_torch(out, shape, dtype)


def tview(c, arr):
    """torch view of a B200Array with the axes reversed (column-major array == reversed row-major view)."""
    torch = c.torch
    real = torch.float64 if arr.dtype in (np.dtype(np.complex128), np.dtype(np.float64)) else torch.float32
    n = arr.size
    flat = arr._owner[: arr.nbytes].view(real)
    if arr.dtype.kind == "c":
        flat = torch.view_as_complex(flat.view(n, 2))
    return flat.view(*reversed(arr.shape)) if arr.shape else flat.view(())


def take(c, arr, labels, restrict):
    """Host copy (column-major axis order) of `arr` with the labels in `restrict` cut down to the given index lists."""
    v = tview(c, arr)
    nd = len(labels)
    for pos, lab in enumerate(labels):
        if lab in restrict:
            idx = c.torch.as_tensor(list(restrict[lab]), device=v.device)
            v = v.index_select(nd - 1 - pos, idx)
    return np.asfortranarray(v.contiguous().cpu().numpy().transpose())


def oracle_slab(ia, ib, ic, a, b):
    """FP64 oracle on host slabs; labels restricted to ONE value that are shared by all three tensors (batch) are squeezed
    so that BackendBase's restatement applies (it rejects hyperindices, binary_einsum.jl:82-83)."""
    from oracle import binary_einsum_base
    ia, ib, ic = list(ia), list(ib), list(ic)
    for lab in [x for x in ia if x in ib and x in ic]:
        pa, pb = ia.index(lab), ib.index(lab)
        assert a.shape[pa] == 1 and b.shape[pb] == 1, "batch labels must be restricted to one value for the oracle"
        a = a.reshape(a.shape[:pa] + a.shape[pa + 1:], order="F"); ia.pop(pa)
        b = b.reshape(b.shape[:pb] + b.shape[pb + 1:], order="F"); ib.pop(pb)
        ic.remove(lab)
    wide = np.complex128 if a.dtype.kind == "c" or b.dtype.kind == "c" else np.float64
    return binary_einsum_base(ic, a.astype(wide), ia, b.astype(wide), ib), ic


def parity_slab(c, ia, ib, ic, A, B, Cc, restrict):
    """rel. Frobenius error of a slab of the device result against the oracle on the same operand slabs."""
    a = take(c, A.data, ia, restrict)
    b = take(c, B.data, ib, restrict)
    got = take(c, Cc.data, ic, restrict)
    ref, ic2 = oracle_slab(ia, ib, ic, a, b)
    got = got.reshape(ref.shape, order="F").astype(ref.dtype)
    return float(np.linalg.norm((got - ref).ravel()) / max(np.linalg.norm(ref.ravel()), 1e-300))


def barrier(c):
    c.torch.cuda.synchronize()
    if c.world > 1:
        c.dist.barrier()
    c.torch.cuda.synchronize()


def max_over_ranks(c, x):
    if c.world > 1:
        t = c.torch.tensor([x], device=c.dev, dtype=c.torch.float64)
        c.dist.all_reduce(t, op=c.dist.ReduceOp.MAX)
        return float(t.item())
    return float(x)


def timed(c, fn, iters, warmup=3, min_ms=0.0):
    """CUDA events on the launching (current torch) stream, warm-up first, barrier + synchronize on both sides, max over ranks.
    min_ms > 0: at least that much device time (sustained figure, long enough for the 50 ms clock / power samples)."""
    torch = c.torch
    for _ in range(warmup):
        fn()
    barrier(c)
    if min_ms > 0:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            fn()
        e1.record()
        torch.cuda.synchronize()
        est = max_over_ranks(c, e0.elapsed_time(e1) / 3)
        iters = int(min(max(iters, min_ms / max(est, 1e-4)), 20000))
        barrier(c)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    barrier(c)
    c.last_iters = iters
    return max_over_ranks(c, ms)


# ---- the headline: config 4b, strong scaling over the free label i -----------------------------------------------------
def headline(c, args, sampler):
    import muscle_b200 as mb
    from muscle_b200 import Index, Tensor, _lib, binary_einsum
    torch = c.torch
    I = lambda s: [Index(x) for x in s]
    ext = dict(CFG4B["ext"])
    if args.cfg4_small:
        ext = {k: 16 for k in ext}
    ia, ib, ic = CFG4B["ia"], CFG4B["ib"], CFG4B["ic"]
    # the cut comes from the library's shard planner (mode ids = positions in the label string "abcdefghi")
    lab = "abcdefghi"
    info = _lib.shard_plan([lab.index(x) for x in ic], [lab.index(x) for x in ia], [ext[x] for x in ia],
                           [lab.index(x) for x in ib], [ext[x] for x in ib], c.world, c.rank)
    if c.world > 1:
        assert info.kind == _lib.SHARD_FREE and lab[info.mode] == "i", (info.kind, info.mode)
        lo, hi = int(info.begin), int(info.end)
    else:
        lo, hi = 0, ext["i"]
    ext_loc = dict(ext, i=hi - lo)
    flops = 8.0 * float(np.prod([ext[x] for x in ext], dtype=np.float64))          # whole job
    h = mb.Handle.get(c.local)

    A = Tensor(dev_rand(c, [ext[x] for x in ia], "complex128", 4000), I(ia))        # replicated (same seed on every rank)
    # B[f,g,d,h,e,i]: i is the slowest mode -> a slab is a contiguous chunk; one seed per i so the global B is the same at every N
    per_i = int(np.prod([ext[x] for x in ib[:-1]]))
    Bflat = torch.empty(2 * per_i * (hi - lo), dtype=torch.float64, device=c.dev)
    for k, i in enumerate(range(lo, hi)):
        g = torch.Generator(device=c.dev); g.manual_seed(41000 + i)
        Bflat[2 * per_i * k: 2 * per_i * (k + 1)].uniform_(-1, 1, generator=g)
    from muscle_b200 import B200Array
    B = Tensor(B200Array.from_torch(Bflat, [ext_loc[x] for x in ib], "complex128"), I(ib))

    step = lambda: binary_einsum(A, B, out=I(ic))
    W_, K_ = max(args.warmup, 3), args.steps
    for _ in range(W_):
        Cc = step()
    barrier(c)
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(K_ + 1)]
    m0 = sampler.mark()
    h.reset_stats()
    barrier(c)
    t0 = time.perf_counter()
    ev[0].record()
    for k in range(K_):
        Cc = step()
        ev[k + 1].record()
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    stats = h.stats()
    m1 = sampler.mark()
    total_ms = ev[0].elapsed_time(ev[K_])
    per_launch = [ev[k].elapsed_time(ev[k + 1]) for k in range(K_)]
    barrier(c)
    ms_per_step = max_over_ranks(c, total_ms) / K_
    value = flops / (ms_per_step * 1e-3) / 1e12
    clocks = sampler.window(m0, m1)

    # parity of this rank's slab against the oracle (outside the timed region): 2 values of a x 2 values of g x the slab's first i
    rng = np.random.default_rng(4102)
    restrict = {"a": sorted(rng.choice(ext["a"], 2, replace=False).tolist()),
                "g": sorted(rng.choice(ext["g"], 2, replace=False).tolist()), "i": [0]}
    par = parity_slab(c, ia, ib, ic, A, B, Cc, restrict)
    pars = [par]
    if c.world > 1:
        pars = [None] * c.world
        c.dist.all_gather_object(pars, par)

    roof = fp64_roofline("gett_kernel<CoreZ<128,64,32,32,32,2>>: ComplexF64 4M on DMMA.8x8x4, operand tiles gathered from the "
                         "native layouts, permuting epilogue (one launch per step)",
                         flops / c.world, float(np.mean(per_launch)), "gett_z_cfg4b_dram_bytes_per_launch")
    roof["share_of_step"] = 1.0
    if c.world > 1:
        roof["traffic"] = None      # the ncu capture is of the N = 1 launch (138 GB: profiles/ncu_r02_summary.md #0); a slab launch was not captured
    else:
        roof["traffic_note"] = ("10.7x the algorithmic bytes at 1.7 % of DRAM bandwidth: 221 waves x (8 A panels + 18.5 B panels = 578 MB) "
                                "that no 126 MB L2 keeps from one wave to the next (DESIGN 3.1)")
    roof["algorithmic_bytes_per_launch"] = 16.0 * (np.prod([ext[x] for x in ia]) + np.prod([ext_loc[x] for x in ib])
                                                   + np.prod([ext_loc[x] for x in ic]))
    out = dict(value=value, ms_per_step=ms_per_step, flops=flops, clocks=clocks, roofline=roof, stats=stats,
               wall_ms_per_step=wall / K_ * 1e3, parity={"rel_frobenius_slab_per_rank": pars, "tolerance": 1e-12,
                                                          "slab": f"a in {restrict['a']}, g in {restrict['g']}, first i of the rank's slab"},
               shard={"label": "i", "range": [lo, hi], "kind": "free index, no collective" if c.world > 1 else "single GPU"})
    return out, (A, B, ext, ext_loc, lo, hi)


def headline_e2e(c, args, state):
    """The same step through the public API with HOST buffers: every step uploads its operands from pinned host memory and
    downloads its result. N = 1: A and B whole. N > 1: every rank uploads 1/N of A (a contiguous chunk of the slowest mode f)
    and NCCL all-gathers it over NVLink (a host would otherwise push N copies of A through PCIe), plus its slab of B, and
    downloads its slab of C. Steps are pipelined over upload / contract / download streams, all copies inside the timed region."""
    import muscle_b200 as mb
    from muscle_b200 import B200Array, Index, Tensor, binary_einsum
    torch, dist = c.torch, c.dist
    I = lambda s: [Index(x) for x in s]
    A, B, ext, ext_loc, lo, hi = state
    ia, ib, ic = CFG4B["ia"], CFG4B["ib"], CFG4B["ic"]
    nA, nB = A.data.size, B.data.size
    nC = int(np.prod([ext_loc[x] for x in ic]))
    partA = nA // c.world                                  # f (slowest mode of A, extent 16) splits evenly for N in 1,2,4,8
    assert partA * c.world == nA
    pinA = torch.empty(2 * partA, dtype=torch.float64).pin_memory()
    pinB = torch.empty(2 * nB, dtype=torch.float64).pin_memory()
    pinC = [torch.empty(2 * nC, dtype=torch.float64).pin_memory() for _ in range(2)]
    # the host copies hold the same synthetic data as the device-resident run
    flatA = A.data._owner[: A.data.nbytes].view(torch.float64)
    pinA.copy_(flatA[2 * partA * c.rank: 2 * partA * (c.rank + 1)])
    pinB.copy_(B.data._owner[: B.data.nbytes].view(torch.float64))
    torch.cuda.synchronize()
    ref_C = binary_einsum(A, B, out=I(ic))
    ref_host = ref_C.data.to_host().reshape(-1, order="F").view(np.float64).copy()
    del ref_C
    h2d = (2 * partA + 2 * nB) * 8
    d2h = 2 * nC * 8
    s_in, s_comp, s_out = torch.cuda.Stream(c.local), torch.cuda.Stream(c.local), torch.cuda.Stream(c.local)
    shapeA, shapeB = [ext[x] for x in ia], [ext_loc[x] for x in ib]

    def one(k, keep, depth=2):
        if len(keep) >= depth:
            keep[0][2].synchronize()
            keep.pop(0)
        with torch.cuda.stream(s_in):
            dA = torch.empty(2 * nA, dtype=torch.float64, device=c.dev)
            dB = torch.empty(2 * nB, dtype=torch.float64, device=c.dev)
            if c.world > 1:
                mine = dA[2 * partA * c.rank: 2 * partA * (c.rank + 1)]
                mine.copy_(pinA, non_blocking=True)
                dist.all_gather_into_tensor(dA, mine)
            else:
                dA.copy_(pinA, non_blocking=True)
            dB.copy_(pinB, non_blocking=True)
            ev_in = torch.cuda.Event(); ev_in.record(s_in)
        with torch.cuda.stream(s_comp):
            s_comp.wait_event(ev_in)
            tA = Tensor(B200Array.from_torch(dA, shapeA, "complex128"), I(ia))
            tB = Tensor(B200Array.from_torch(dB, shapeB, "complex128"), I(ib))
            z = binary_einsum(tA, tB, out=I(ic))
            ev_c = torch.cuda.Event(); ev_c.record(s_comp)
        with torch.cuda.stream(s_out):
            s_out.wait_event(ev_c)
            pinC[k & 1].copy_(z.data._owner[: z.data.nbytes].view(torch.float64), non_blocking=True)
            ev_o = torch.cuda.Event(); ev_o.record(s_out)
        for t_, s_ in ((dA, s_comp), (dB, s_comp), (z.data._owner, s_out)):
            t_.record_stream(s_)
        keep.append((dA, z, ev_o))

    def run(n):
        keep = []
        for k in range(n):
            one(k, keep)
        for st in (s_in, s_comp, s_out):
            st.synchronize()

    run(2)                                                  # warm-up (allocator, NCCL channel set-up)
    check = float(np.abs(pinC[1].numpy() - ref_host).max())
    barrier(c)
    t0 = time.perf_counter()
    run(1)
    serial_ms = max_over_ranks(c, (time.perf_counter() - t0) * 1e3)
    Ke = max(3, min(args.steps, 8))
    barrier(c)
    t0 = time.perf_counter()
    run(Ke)
    e2e_ms = max_over_ranks(c, (time.perf_counter() - t0) * 1e3) / Ke
    flops = 8.0 * float(np.prod([ext[x] for x in ext], dtype=np.float64))
    return {"value": flops / (e2e_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "h2d_bytes_per_step": h2d * c.world,
            "d2h_bytes_per_step": d2h * c.world, "ms_per_step": e2e_ms, "steps": Ke,
            "serial_ms_per_step": serial_ms, "serial_value": flops / (serial_ms * 1e-3) / 1e12,
            "max_abs_diff_vs_device_resident_result": check,
            "nvlink_allgather_bytes_per_step": (16 * nA if c.world > 1 else 0),
            "timing": "host wall clock around Ke pipelined steps, all streams synchronised on both sides, max over ranks",
            "api": "pinned host buffers -> H2D (N > 1: 1/N of A per rank + NCCL all_gather over NVLink) -> binary_einsum -> D2H of "
                   "the rank's slab of C; steps pipelined over upload/contract/download streams (serial_* = one un-pipelined step); "
                   "h2d/d2h bytes are the whole job's (all ranks)"}


# ---- per-config block ----------------------------------------------------------------------------------------------------
def per_config_n1(c, args, sampler):
    """BASELINE configs 1, 2, 3, 4a, 5 on one GPU: >= 20 timed iterations each after >= 3 warm-ups, device-resident operands."""
    import muscle_b200 as mb
    from muscle_b200 import Index, Tensor, binary_einsum
    torch = c.torch
    I = lambda s: [Index(x) for x in s]
    h = mb.Handle.get(c.local)
    iters = max(20, min(args.steps, 50))
    out = {}

    def run_one(name, dtype, ext, ia, ib, ic, seeds, restrict, roof_fn, tol, note):
        A = Tensor(dev_rand(c, [ext[x] for x in ia], dtype, seeds[0]), I(ia))
        B = Tensor(dev_rand(c, [ext[x] for x in ib], dtype, seeds[1]), I(ib))
        fn = lambda: binary_einsum(A, B, out=I(ic))
        h.reset_stats()
        Cc = fn()
        s1 = h.stats()
        m0 = sampler.mark()
        ms = timed(c, fn, iters, min_ms=400.0)
        m1 = sampler.mark()
        labels = set(ia) | set(ib)
        cplx = dtype.startswith("complex")
        flops = (8.0 if cplx else 2.0) * float(np.prod([ext[x] for x in labels], dtype=np.float64))
        par = parity_slab(c, ia, ib, ic, A, B, Cc, restrict)
        out[name] = {"workload": note, "dtype": dtype, "value": flops / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms": ms,
                     "iters": c.last_iters, "flops": flops, "roofline": roof_fn(flops, ms), "clocks": sampler.window(m0, m1),
                     "launches_per_call": {k: v for k, v in s1.items() if v and k.startswith("launches")},
                     "parity": {"rel_frobenius": par, "tolerance": tol, "slab": {k: list(v) for k, v in restrict.items()}}}
        del A, B, Cc
        torch.cuda.empty_cache()

    n = 64
    run_one("config1_rank4_dim64", "complex128", dict(i=n, j=n, k=n, l=n, m=n, n=n), "kilj", "nlmk", "mjni", (1000, 1001),
            {"i": [5, 40], "n": [3, 33]},
            lambda f, ms: fp64_roofline("gett_kernel<CoreZ<128,64,...>>", f, ms, "gett_z_cfg1_dram_bytes_per_launch"), 1e-12,
            "BASELINE configs[0]: two random ComplexF64 rank-4 tensors, dim 64, two summed labels (4096^3), scrambled layout "
            "A[k,i,l,j] B[n,l,m,k] -> C[m,j,n,i]")
    out["config2_mps_mpo_chain"] = config2_chain(c, args, sampler, max(iters, 50))
    chi, Dd, beta = 256, 8, 8
    run_one("config3_peps_batched_c64", "complex64", dict(l=chi, k=Dd, b=Dd, m=chi, q=Dd, r=chi, z=beta), "lkbmz", "mkqrz", "lbqrz",
            (3000, 3100), {"l": list(range(0, 256, 8)), "z": [5]},
            lambda f, ms: tf32_roofline("2 x K1 split-writer pack + tf32_gemm_kernel (tcgen05/TMEM, TF32 + BF16 split, CTA pairs)", f, ms,
                                        "cfg3_dram_bytes_per_call"), 1e-5,
            "BASELINE configs[2]: PEPS double-layer ComplexF32 D=8 chi=256 with batch hyperindex beta=8 (2048^3 x 8)")
    n = 16
    run_one("config4a_rank6_dim16", "complex128", {x: n for x in "abcdefghi"}, "adbecf", "fgdhei", "abcghi", (4000, 4001),
            {"a": [1, 9], "g": [4, 12]},
            lambda f, ms: fp64_roofline("gett_kernel<CoreZ<128,64,...>>", f, ms), 1e-12,
            "BASELINE configs[3], literal reading: rank-6 ComplexF64, dim 16 per label (4096^3)")
    n = 8
    run_one("config5_rank8_c64", "complex64", {x: n for x in "abcdefghpqrs"}, "aebfcgdh", "hpgqfres", "srqpdcba", (5000, 5100),
            {"a": [2], "s": [6]},
            lambda f, ms: tf32_roofline("2 x K1 split-writer pack + tf32_gemm_kernel", f, ms, "cfg5_dram_bytes_per_call"), 1e-5,
            "BASELINE configs[4] on one GPU: rank-8 ComplexF32, dim 8, 4 summed labels (4096^3), interleaved labels, reversed output")
    return out


def config2_chain(c, args, sampler, iters):
    """configs[1]: MPS-MPO transfer contraction, ComplexF64, chi=1024, d=2, w=8 - the three-contraction chain
    T = E[a,w,b] A[b,s,c]; T' = T W[w,s,t,v] -> [a,t,v,c]; E' = T' conj-site[a,t,e] -> [e,v,c]."""
    from muscle_b200 import Index, Tensor, binary_einsum
    torch = c.torch
    I = lambda s: [Index(x) for x in s]
    ext = dict(a=CHI, b=CHI, c=CHI, e=CHI, w=W, v=W, s=D, t=D)
    steps = [("2a", "awb", "bsc", "awsc"), ("2b", "awsc", "wstv", "atvc"), ("2c", "atvc", "ate", "evc")]
    flops = [8.0 * float(np.prod([ext[x] for x in set(ia) | set(ib)], dtype=np.float64)) for _, ia, ib, _ in steps]
    T = {"E": Tensor(dev_rand(c, [CHI, W, CHI], "complex128", 2000), I("awb")),
         "A": Tensor(dev_rand(c, [CHI, D, CHI], "complex128", 2001), I("bsc")),
         "W": Tensor(dev_rand(c, [W, D, D, W], "complex128", 2002), I("wstv")),
         "Ab": Tensor(dev_rand(c, [CHI, D, CHI], "complex128", 2003), I("ate"))}
    for _ in range(3):
        x = binary_einsum(T["E"], T["A"], out=I("awsc"))
        y = binary_einsum(x, T["W"], out=I("atvc"))
        z = binary_einsum(y, T["Ab"], out=I("evc"))
    barrier(c)
    ev = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(iters)]
    m0 = sampler.mark()
    for k in range(iters):
        ev[k][0].record()
        x = binary_einsum(T["E"], T["A"], out=I("awsc"))
        ev[k][1].record()
        y = binary_einsum(x, T["W"], out=I("atvc"))
        ev[k][2].record()
        z = binary_einsum(y, T["Ab"], out=I("evc"))
        ev[k][3].record()
    torch.cuda.synchronize()
    m1 = sampler.mark()
    total = ev[0][0].elapsed_time(ev[iters - 1][3]) / iters
    t = [float(np.mean([ev[k][j].elapsed_time(ev[k][j + 1]) for k in range(iters)])) for j in range(3)]
    # parity per step on slabs (each step checked against the oracle fed with the DEVICE's own input of that step)
    par = {"2a": parity_slab(c, "awb", "bsc", "awsc", T["E"], T["A"], x, {"a": [7, 500], "c": list(range(0, 1024, 64))}),
           "2b": parity_slab(c, "awsc", "wstv", "atvc", x, T["W"], y, {"a": [3, 900], "c": list(range(5, 1024, 64))}),
           "2c": parity_slab(c, "atvc", "ate", "evc", y, T["Ab"], z, {"e": [11, 333], "c": list(range(9, 1024, 64))})}
    bytes_2b = 16.0 * (CHI * W * D * CHI + W * D * D * W + CHI * D * W * CHI)
    return {"workload": f"BASELINE configs[1]: MPS-MPO transfer contraction ComplexF64 chi={CHI} d={D} w={W}; 3-contraction chain 2a,2b,2c",
            "dtype": "complex128", "value": sum(flops) / (total * 1e-3) / 1e12, "unit": "TFLOP/s", "ms": total, "iters": iters,
            "flops": sum(flops), "step_breakdown_ms": {"2a": t[0], "2b": t[1], "2c": t[2]},
            "roofline": fp64_roofline("gett_kernel<CoreZ<128,64,32,32,32,2>> (steps 2a, 2c)", 0.5 * (flops[0] + flops[2]), 0.5 * (t[0] + t[2]),
                                      "gett_z_128x64_dram_bytes_per_launch"),
            "roofline_2b": hbm_roofline("stream_kernel<CoreZ<64,16,16,16,8,2>> (step 2b, N=K=16: persistent, B resident in smem)", bytes_2b, t[1],
                                        "stream_2b_dram_bytes_per_launch"),
            "clocks": sampler.window(m0, m1), "parity": {"rel_frobenius_per_step": par, "tolerance": 1e-12}}


def k1_rooflines(c, args, sampler):
    """K1 permute / matricise kernels against the measured HBM copy peak: algorithmic bytes = 2 * sizeof(T) * numel."""
    from muscle_b200 import Index, Tensor
    I = lambda s: [Index(x) for x in s]
    out = {}
    cases = [("c128_64^4_kilj_to_ijkl (config-1 A pack)", "complex128", (64, 64, 64, 64), "kilj", "ijkl"),
             ("c128_4096^2_transpose", "complex128", (4096, 4096), "ij", "ji"),
             ("c64_(256,8,8,256,8)_lkbmz_to_mklbz (config-3 A matricise)", "complex64", (256, 8, 8, 256, 8), "lkbmz", "mklbz"),
             ("f32_8192^2_transpose", "float32", (8192, 8192), "ij", "ji")]
    for name, dt, shape, src, dst in cases:
        t = Tensor(dev_rand(c, shape, dt, 77), I(src))
        fn = lambda: t.permutedims(I(dst))
        m0 = sampler.mark()
        ms = timed(c, fn, 20, min_ms=200.0)
        nbytes = 2.0 * t.data.nbytes
        out[name] = hbm_roofline("permute kernels (csrc/permute.cu)", nbytes, ms)
        out[name]["clocks"] = sampler.window(m0, sampler.mark())
        out[name]["iters"] = c.last_iters
        del t
        c.torch.cuda.empty_cache()
    return out


def latency_probe(c):
    """Launch-bound tiny contraction (the size of every case in the reference's own test-suite): C-ABI call and Python front-end."""
    import ctypes as C
    from muscle_b200 import Index, Tensor, _lib, binary_einsum
    I = lambda s: [Index(x) for x in s]
    A = Tensor(np.ones((2, 3)), I("ij")).to_device(c.local)
    B = Tensor(np.ones((3, 4)), I("jk")).to_device(c.local)
    Cc = binary_einsum(A, B)
    h = _lib.Handle.get(c.local)
    L = _lib.lib()
    mc, ma, mb_ = _lib.i32([0, 2]), _lib.i32([0, 1]), _lib.i32([1, 2])
    ea, eb = _lib.i64([2, 3]), _lib.i64([3, 4])
    args = (h.ptr, C.c_void_p(Cc.data.ptr), _lib.F64, 2, mc, None, C.c_void_p(A.data.ptr), _lib.F64, 2, ma, ea, None,
            C.c_void_p(B.data.ptr), _lib.F64, 2, mb_, eb, None)
    n = 3000
    for _ in range(200):
        L.mb200_binary_einsum(*args)
    c.torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        L.mb200_binary_einsum(*args)
    c.torch.cuda.synchronize()
    abi = (time.perf_counter() - t0) / n * 1e6
    for _ in range(200):
        binary_einsum(A, B)
    c.torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        binary_einsum(A, B)
    c.torch.cuda.synchronize()
    py = (time.perf_counter() - t0) / n * 1e6
    return abi, py


# ---- sharded configs at N > 1 ------------------------------------------------------------------------------------------------
def sharded_configs(c, args, sampler):
    """config 3 sharded over the batch label (no collective) and config 5 sliced over a summed label (partial C add-reduced:
    NCCL all_reduce baseline, fused peer-memory reduce-scatter, fused all-reduce). STRONG scaling; times are CUDA events, max
    over ranks; parity of rank 0's result against the oracle outside the timed region."""
    import muscle_b200 as mb
    from muscle_b200 import Index, Tensor, binary_einsum
    from muscle_b200 import dist as mdist
    torch, dist = c.torch, c.dist
    I = lambda s: [Index(x) for x in s]
    out = {}
    iters = 20
    world, rank = c.world, c.rank

    # config 3 over beta
    chi, Dd, beta = 256, 8, 8
    bl = max(1, beta // world)
    z0 = rank * bl
    A3 = Tensor(dev_rand(c, [chi, Dd, Dd, chi, bl], "complex64", 3000 + z0), I("lkbmz"))
    B3 = Tensor(dev_rand(c, [chi, Dd, Dd, chi, bl], "complex64", 3100 + z0), I("mkqrz"))
    fn3 = lambda: binary_einsum(A3, B3, out=I("lbqrz"))
    C3 = fn3()
    m0 = sampler.mark()
    ms = timed(c, fn3, iters, min_ms=300.0)
    flops3 = 8.0 * float(chi * Dd) ** 3 * (bl * min(world, beta))
    par = parity_slab(c, "lkbmz", "mkqrz", "lbqrz", A3, B3, C3, {"l": list(range(0, 256, 8)), "z": [0]})
    pars = [None] * world
    dist.all_gather_object(pars, par)
    out["config3_peps_batched_c64_over_beta"] = {
        "workload": f"PEPS double-layer ComplexF32 D=8 chi=256 beta=8 (2048^3 x 8), batch label sharded {min(world, beta)}x, no collective",
        "scaling": "strong", "value": flops3 / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms": ms, "iters": iters, "flops": flops3,
        "clocks": sampler.window(m0, sampler.mark()),
        "parity": {"rel_frobenius_slab_per_rank": pars, "tolerance": 1e-5}}
    del A3, B3, C3
    torch.cuda.empty_cache()

    # config 5: summed label h sliced over the ranks
    n = 8
    ia, ib, ic = "aebfcgdh", "hpgqfres", "srqpdcba"
    hl = max(1, n // world)
    A5 = Tensor(dev_rand(c, [n] * 7 + [hl], "complex64", 5000 + rank), I(ia))
    B5 = Tensor(dev_rand(c, [hl] + [n] * 7, "complex64", 5100 + rank), I(ib))
    flops5 = 8.0 * float(n ** 12) * (hl * min(world, n) / n)
    restrict = {"a": [2], "s": [6]}
    # the oracle needs every rank's operand slabs: gather the (small) restricted slabs on every rank
    a_s = take(c, A5.data, ia, restrict)
    b_s = take(c, B5.data, ib, restrict)
    parts = [None] * world
    dist.all_gather_object(parts, (a_s, b_s))
    ref = None
    for pa, pb in parts:
        r, _ = oracle_slab(ia, ib, ic, pa, pb)
        ref = r if ref is None else ref + r

    def err_full(Ct):
        got = take(c, Ct.data, ic, restrict).reshape(ref.shape, order="F").astype(np.complex128)
        return float(np.linalg.norm((got - ref).ravel()) / np.linalg.norm(ref.ravel()))

    res5 = {"workload": f"rank-8 ComplexF32 dim 8, 4 summed (4096^3); summed label h sliced {min(world, n)}x, partial C (134 MB) add-reduced",
            "scaling": "strong", "flops": flops5, "allreduce_bytes": 8 * n ** 8, "variants": {}}

    def nccl_step():
        cc = binary_einsum(A5, B5, out=I(ic))
        mdist.all_reduce_sum(cc)
        return cc
    m0 = sampler.mark()
    ms = timed(c, nccl_step, iters, min_ms=300.0)
    ms_gemm = timed(c, lambda: binary_einsum(A5, B5, out=I(ic)), iters, min_ms=100.0)
    res5["variants"]["nccl_all_reduce"] = {"value": flops5 / (ms * 1e-3) / 1e12, "unit": "TFLOP/s", "ms": ms, "ms_contraction_only": ms_gemm,
                                           "semantics": "all-reduce (every rank ends with the full C)",
                                           "parity_rel_frobenius": err_full(nccl_step()), "clocks": sampler.window(m0, sampler.mark())}
    # fused peer-memory variants
    for key, fname, sem in (("fused_all_reduce", "sum_slice_all_reduce", "all-reduce (every rank ends with the full C)"),
                            ("fused_reduce_scatter", "sum_slice_reduce_scatter", "reduce-scatter (each rank ends with its 1/N slab of C)")):
        f = getattr(mdist, fname, None)
        if f is None:
            continue
        try:
            got = f(A5, B5, I(ic))
            if key == "fused_all_reduce":
                err = err_full(got)
            else:
                # slab = flat column-major range [rank*slab, (rank+1)*slab): compare with the NCCL all-reduced result's range
                full = nccl_step().data.to_host().reshape(-1, order="F")
                mine = got.data.to_host().reshape(-1, order="F")
                sl = full[rank * mine.size:(rank + 1) * mine.size]
                err = float(np.linalg.norm(mine - sl) / max(np.linalg.norm(sl), 1e-30))
            m0 = sampler.mark()
            ms_f = timed(c, lambda: f(A5, B5, I(ic)), iters, min_ms=300.0)
            res5["variants"][key] = {"value": flops5 / (ms_f * 1e-3) / 1e12, "unit": "TFLOP/s", "ms": ms_f, "semantics": sem,
                                     ("parity_rel_frobenius" if key == "fused_all_reduce" else "rel_err_vs_nccl_all_reduce"): err,
                                     "clocks": sampler.window(m0, sampler.mark())}
        except Exception as e:  # noqa: BLE001
            res5["variants"][key] = {"error": repr(e)[:400]}
    best = max((v for v in res5["variants"].values() if "value" in v and v.get("semantics", "").startswith("all-reduce")),
               key=lambda v: v["value"])
    res5["value"], res5["ms"], res5["unit"] = best["value"], best["ms"], "TFLOP/s"
    res5["parity"] = {"tolerance": 1e-5, "slab": restrict}
    try:
        kinds = mdist.allreduce_plumbing_info()
        res5["fused_all_reduce_plumbing"] = {"peer_buffers": kinds[0][0], "nvls_multicast": kinds[0][1]} if kinds else None
    except Exception:  # noqa: BLE001
        pass
    out["config5_summed_slice"] = res5
    del A5, B5
    torch.cuda.empty_cache()
    return out


# ------------------------------------------------------------------------------------------------
def run_ours(args):
    c = make_ctx()
    torch, dist = c.torch, c.dist
    # stdout carries exactly one JSON line: NCCL's own banner / debug output goes to stderr
    os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")
    torch.cuda.set_device(c.local)
    if c.world > 1:
        dist.init_process_group("nccl", device_id=torch.device(c.dev))
    import muscle_b200 as mb
    mb.Handle.get(c.local)

    sampler = ClockSampler(c.local)
    sampler.start()                       # returns after nvidia-smi has written its first samples
    if c.world > 1:
        dist.barrier()

    head, state = headline(c, args, sampler)
    e2e = None
    if not args.skip_e2e:
        e2e = headline_e2e(c, args, state)
    del state
    torch.cuda.empty_cache()

    per_config, k1, lat = None, None, None
    if not args.skip_configs:
        if c.world == 1:
            per_config = per_config_n1(c, args, sampler)
            k1 = k1_rooflines(c, args, sampler)
            lat = latency_probe(c)
        else:
            per_config = sharded_configs(c, args, sampler)

    cpu = None
    if c.world == 1 and not args.skip_cpu:
        tf, mean, sample, cores = time_cpu_cfg4b(3, 1, budget_s=30.0)
        cpu = {"value": tf, "unit": "TFLOP/s", "cores": cores, "kind": "port", "sample": sample, "ms_per_step": mean * 1e3,
               "what": "CPU restatement of Muscle BackendBase (numpy permutedims copies + OpenBLAS zgemm); Julia unavailable"}
    sampler.stop()

    if c.rank == 0:
        stats = head["stats"]
        line = {
            "metric": METRIC, "value": head["value"], "unit": "TFLOP/s", "n_gpus": c.world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": head["ms_per_step"], "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD,
                       "arithmetic": "ComplexF64 as 4M real products on FP64 tensor cores (DMMA.8x8x4), FP64 accumulation",
                       "flops_per_step": head["flops"],
                       "l2": "no flush needed: every step streams 8.6 GB of operands and writes 4.3 GB (>> 126 MB L2)",
                       "parallelism": (f"free label i of B and C cut into {c.world} slabs (mb200_shard_plan), A replicated, no collective"
                                       if c.world > 1 else "single GPU"),
                       "shard": head["shard"]},
            "pct_of_fp64_tensor_peak": 100.0 * head["value"] / c.world / FP64_TENSOR_PEAK_TFLOPS,
            "wall_ms_per_step": head["wall_ms_per_step"],
            "roofline": head["roofline"], "parity": head["parity"], "e2e": e2e, "cpu_baseline": cpu,
            "gpu_launches": int(stats["launches_total"]),
            "gpu_launch_breakdown": {k: v for k, v in stats.items() if v},
            "clocks": head["clocks"],
            "measured_library_peaks": {"source": "profiles/peaks_r01.json (tools/measure_peaks.py, cuBLAS on this pool)", **{
                k: (v.get("burst_tflops") if isinstance(v, dict) else v) for k, v in CUBLAS.items() if k != "gpu"}},
            "per_config": per_config, "roofline_k1": k1,
        }
        if lat:
            line["tiny_contraction_latency_us"] = {"c_abi_call": lat[0], "python_front_end_call": lat[1]}
        emit(line)
    if c.world > 1:
        dist.barrier()
        dist.destroy_process_group()


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
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--skip-cpu", action="store_true")
    ap.add_argument("--skip-e2e", action="store_true")
    ap.add_argument("--skip-configs", action="store_true", help="headline only (no per_config / roofline_k1 blocks)")
    ap.add_argument("--cfg4-small", action="store_true", help="headline at dim 16 (4096^3) instead of 16384^3 (debugging)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
