"""A/B of the tcgen05 path's two split schemes (tf32.cu): mixed TF32 + BF16 (default, 8 MMAs per 8 k of a complex
product) against 3xTF32 (MB200_SPLIT_SCHEME=3xtf32, 12 MMAs). Per scheme, in its own process (the scheme is read once):
relative Frobenius error against an fp64 product at K = 256 .. 16384 (ComplexF32, Float32; uniform and normal
inputs) and the timing of configs 3 / 5 and an 8192^3 Float32 contraction. Writes gpurun_out/split_scheme.json.

    python tools/probe_split_scheme.py            # parent: runs both schemes
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))


def child():
    import numpy as np
    import torch

    import bench_kernels as bk
    import muscle_b200 as mb
    from muscle_b200 import _lib

    sys.path.insert(0, os.path.join(ROOT, "tests"))
    out = {"scheme": os.environ.get("MB200_SPLIT_SCHEME", "mixed"), "accuracy": [], "timing": []}
    h = _lib.Handle.get()

    I = lambda t: [mb.Index(c) for c in t]

    def contract(a, ia, b, ib, ic):
        h.set_path(mb.PATH_TCGEN05_TF32)
        try:
            h.reset_stats()
            c = mb.binary_einsum(mb.BackendB200(), I(ic), mb.Tensor(a, I(ia)).to_device(), mb.Tensor(b, I(ib)).to_device())
            assert h.stats()["launches_tcgen05"] == 1, h.stats()
            return c.to_host().data
        finally:
            h.set_path(mb.PATH_AUTO)

    rng = np.random.default_rng(5)
    for dist in ("uniform", "normal"):
        for dt, wide in (("complex64", np.complex128), ("float32", np.float64)):
            for K in (256, 4096, 16384):
                def gen(shape):
                    f = (lambda s: rng.uniform(-1, 1, s)) if dist == "uniform" else rng.standard_normal
                    x = f(shape)
                    if dt == "complex64":
                        x = x + 1j * f(shape)
                    return np.asfortranarray(x.astype(dt))
                a, b = gen((K, 256)), gen((K, 384))
                ref = a.astype(wide).T @ b.astype(wide)
                got = contract(a, "ki", b, "kj", "ij")
                err = float(np.linalg.norm(got.astype(wide) - ref) / np.linalg.norm(ref))
                out["accuracy"].append({"dtype": dt, "dist": dist, "K": K, "rel_frobenius": err})
                print("ACC %-10s %-8s K=%6d  %.3e" % (dt, dist, K, err), flush=True)
    cases = [("cfg3 PEPS c64", dict(l=256, k=8, b=8, m=256, q=8, r=256, z=8), "lkbmz", "mkqrz", "lbqrz", "complex64"),
             ("cfg5 rank8 c64", {c: 8 for c in "abcdefghpqrs"}, "aebfcgdh", "hpgqfres", "srqpdcba", "complex64"),
             ("c64 4096^3 aligned", dict(i=4096, j=4096, k=4096), "ki", "kj", "ij", "complex64"),
             ("f32 8192^3 aligned", dict(i=8192, j=8192, k=8192), "ki", "kj", "ij", "float32")]
    for name, ext, ia, ib, ic, dt in cases:
        r = bk.einsum_case(name, ext, ia, ib, ic, dt, iters=8)
        out["timing"].append({k: r[k] for k in ("name", "tflops_best", "tflops_mean", "ms_mean") if k in r})
        print("EINSUM %-22s %8.2f TF/s best %8.2f mean %.3f ms" % (name, r["tflops_best"], r["tflops_mean"], r["ms_mean"]), flush=True)
    print("JSON " + json.dumps(out), flush=True)


def main():
    if os.environ.get("MB200_PROBE_CHILD"):
        return child()
    res = []
    for scheme in ("mixed", "3xtf32"):
        env = dict(os.environ, MB200_PROBE_CHILD="1")
        if scheme == "3xtf32":
            env["MB200_SPLIT_SCHEME"] = "3xtf32"
        else:
            env.pop("MB200_SPLIT_SCHEME", None)
        print("==== scheme", scheme, flush=True)
        p = subprocess.run([sys.executable, os.path.abspath(__file__)], env=env, capture_output=True, text=True, timeout=900)
        sys.stdout.write(p.stdout)
        if p.returncode != 0:
            sys.stdout.write(p.stderr[-3000:])
        for line in p.stdout.splitlines():
            if line.startswith("JSON "):
                res.append(json.loads(line[5:]))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    with open(os.path.join(ROOT, "gpurun_out", "split_scheme.json"), "w") as f:
        json.dump(res, f, indent=1)


if __name__ == "__main__":
    main()
