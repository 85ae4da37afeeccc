"""A/B: K1 pure transpositions through the TMA-staged kernel (flags=2) against the register-tile / smem-tile kernels (flags=4).
GB/s = 2 * sizeof(T) * numel / time (CUDA events, 20 iterations after 3 warm-ups), peak = MEASURED_PEAKS.json hbm_gbs."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import muscle_b200 as mb  # noqa: E402
from muscle_b200 import B200Array, Index, Tensor  # noqa: E402

CASES = [
    ("c128 64^4 kilj->ijkl (config-1 A pack)", "complex128", (64, 64, 64, 64), (1, 3, 0, 2)),
    ("c128 64^4 nlmk->lknm (config-1 B pack)", "complex128", (64, 64, 64, 64), (1, 3, 0, 2)[::-1]),
    ("c128 4096^2 transpose", "complex128", (4096, 4096), (1, 0)),
    ("c128 rank-6 16^6 reversal", "complex128", (16,) * 6, (5, 4, 3, 2, 1, 0)),
    ("c64 8192^2 transpose", "complex64", (8192, 8192), (1, 0)),
    ("c64 (256,8,8,256,8) -> (3,1,0,2,4)", "complex64", (256, 8, 8, 256, 8), (3, 1, 0, 2, 4)),
    ("f64 8192^2 transpose", "float64", (8192, 8192), (1, 0)),
    ("f32 8192^2 transpose", "float32", (8192, 8192), (1, 0)),
    ("f32 (1024,64,256) -> (2,1,0)", "float32", (1024, 64, 256), (2, 1, 0)),
    ("c128 (2,1024,1024,8) -> (1,0,3,2) d=2 fastest", "complex128", (2, 1024, 1024, 8), (1, 0, 3, 2)),
]


def main():
    peak = 6540.5
    try:
        peak = float(json.load(open("MEASURED_PEAKS.json"))["hbm_gbs"])
    except Exception:
        pass
    out = []
    for name, dt, shape, perm in CASES:
        n = int(np.prod(shape))
        real = torch.float64 if dt in ("complex128", "float64") else torch.float32
        flat = torch.rand((2 if "complex" in dt else 1) * n, dtype=real, device="cuda:0")
        t = Tensor(B200Array.from_torch(flat, shape, dt), [Index(i) for i in range(len(shape))])
        row = {"case": name}
        for tag, flags in (("tma", 2), ("regt", 4)):
            for _ in range(3):
                t.permutedims(list(perm), flags=flags)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(20):
                r = t.permutedims(list(perm), flags=flags)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / 20
            row[tag + "_gbs"] = 2.0 * t.data.nbytes / (ms * 1e-3) / 1e9
            row[tag + "_frac"] = row[tag + "_gbs"] / peak
        a = t.permutedims(list(perm), flags=2).data.to_host()
        b = t.permutedims(list(perm), flags=4).data.to_host()
        row["bit_equal"] = bool(np.array_equal(a, b))
        out.append(row)
        print("K1 %-48s tma %7.0f GB/s (%.3f)   regT/smem %7.0f GB/s (%.3f)  equal=%s" % (
            name, row["tma_gbs"], row["tma_frac"], row["regt_gbs"], row["regt_frac"], row["bit_equal"]), flush=True)
        del t, flat
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump({"hbm_peak_gbs": peak, "rows": out}, open("gpurun_out/ab_permute_tma.json", "w"), indent=1)


if __name__ == "__main__":
    main()
