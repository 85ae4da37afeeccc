"""Measures the practical FP64 / TF32 ceilings on this box with cuBLAS (through torch.matmul):
DGEMM, ZGEMM, SGEMM (TF32 on and off), CGEMM at 8192^3 / 4096^3 — burst (best of N) and sustained
(back-to-back for ~2 s). Writes gpurun_out/peaks_extra.json. These are yardsticks for the roofline
denominators that MEASURED_PEAKS.json does not carry (it has only HBM copy and bf16)."""
import json
import os
import sys
import time

import torch


def bench(fn, flops, burst_iters=6, sustain_s=2.0):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(burst_iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    n = max(3, int(sustain_s / best))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    sustained = e0.elapsed_time(e1) * 1e-3 / n
    return {"burst_tflops": flops / best / 1e12, "sustained_tflops": flops / sustained / 1e12,
            "burst_ms": best * 1e3, "sustained_ms": sustained * 1e3}


def main():
    out = {"gpu": torch.cuda.get_device_name(0)}
    dev = "cuda:0"
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float64); b = torch.randn(n, n, device=dev, dtype=torch.float64)
    out["dgemm_8192"] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    del a, b
    n = 4096
    a = torch.randn(n, n, device=dev, dtype=torch.complex128); b = torch.randn(n, n, device=dev, dtype=torch.complex128)
    out["zgemm_4096"] = bench(lambda: torch.matmul(a, b), 8.0 * n ** 3)
    del a, b
    n = 8192
    a = torch.randn(n, n, device=dev, dtype=torch.float32); b = torch.randn(n, n, device=dev, dtype=torch.float32)
    torch.backends.cuda.matmul.allow_tf32 = True
    out["sgemm_tf32_8192"] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    torch.backends.cuda.matmul.allow_tf32 = False
    out["sgemm_fp32_8192"] = bench(lambda: torch.matmul(a, b), 2.0 * n ** 3)
    del a, b
    n = 4096
    a = torch.randn(n, n, device=dev, dtype=torch.complex64); b = torch.randn(n, n, device=dev, dtype=torch.complex64)
    out["cgemm_4096"] = bench(lambda: torch.matmul(a, b), 8.0 * n ** 3)
    del a, b
    # HBM copy (same definition as MEASURED_PEAKS.json: read+write bytes)
    x = torch.empty(1 << 30, dtype=torch.bfloat16, device=dev); y = torch.empty_like(x)
    r = bench(lambda: y.copy_(x), 1.0)
    out["hbm_copy_gbs"] = 2 * x.numel() * 2 / (r["burst_ms"] * 1e-3) / 1e9
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/peaks_extra.json", "w"), indent=1)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
