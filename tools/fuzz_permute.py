"""Randomised bit-exactness check of mb200_permute (all kernel families) against numpy.transpose.
    python tools/fuzz_permute.py [ncases] [seed]"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from muscle_b200 import Index, Tensor  # noqa: E402


def main():
    ncases = int(sys.argv[1]) if len(sys.argv) > 1 else 300
    rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
    pool = [1, 2, 2, 3, 4, 4, 5, 6, 8, 8, 12, 16, 20, 32, 33, 64, 100, 128, 260, 1028]
    bad = 0
    for c in range(ncases):
        rank = int(rng.integers(1, 7))
        shape = []
        budget = 1 << 21
        for _ in range(rank):
            e = int(rng.choice([p for p in pool if p <= max(budget, 1)]))
            shape.append(e)
            budget //= e
        perm = list(rng.permutation(rank))
        dt = str(rng.choice(["float32", "float64", "complex64", "complex128"]))
        n = int(np.prod(shape))
        x = np.arange(n, dtype=np.float64).reshape(shape, order="F")
        x = (x + 1j * (x + 0.5)).astype(dt) if dt.startswith("complex") else x.astype(dt)
        x = np.asfortranarray(x)
        t = Tensor(x, [Index(i) for i in range(rank)]).to_device()
        got = t.permutedims([int(p) for p in perm]).to_host().data
        ok = np.array_equal(got, np.transpose(x, perm))
        if not ok:
            bad += 1
            print("MISMATCH", shape, perm, dt, flush=True)
    print(f"fuzz_permute: {ncases} cases, {bad} mismatches")
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
