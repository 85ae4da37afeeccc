"""Generates tests/golden/dangling_golden.npz: contractions with DANGLING labels (carried by one operand only, absent from the output).
BackendBase rejects them (src/Operations/binary_einsum.jl:82-83); the backends BackendB200 replaces - cuTENSOR.contract!
(ext/MuscleCUDAExt.jl:30-38) and OMEinsum (ext/MuscleOMEinsumExt.jl:40-59) - sum them. Outputs come from the plain-C loop nest
(oracle/einsum_ref.c: C += A * B over EVERY label, the explicit loop of test/integration/omeinsum.jl:225-236), accumulated in double
precision and rounded once; the NumPy oracle (CPU test) and the CUDA path (GPU test) are both checked against them.

    python tests/golden/make_golden_dangling.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import random_array  # noqa: E402
from oracle.build_oracle import einsum_loops  # noqa: E402

# name, extents, inds_a, inds_b, inds_c      (x, y: dangling). The last case exceeds 2^20 MACs: the pre-reduce route.
CASES = [
    ("fold_a", dict(i=5, j=7, k=3, x=4), "ixj", "jk", "ik"),
    ("fold_b", dict(i=5, j=7, k=3, y=6), "ij", "yjk", "ki"),
    ("fold_both", dict(i=4, j=5, k=3, x=2, y=3), "xij", "jky", "ik"),
    ("fold_batch", dict(i=4, j=5, k=3, z=2, x=3), "izxj", "jkz", "kiz"),
    ("fold_to_scalar", dict(i=6, x=5), "ix", "i", ""),
    ("prereduce_a", dict(i=16, j=8, k=264, x=32), "ixj", "jk", "ki"),
]
DTYPES = ["float32", "float64", "complex64", "complex128"]


def main():
    out = {}
    for name, ext, ia, ib, ic in CASES:
        for dt in DTYPES:
            rng = np.random.default_rng(4321)
            a = random_array(rng, tuple(ext[c] for c in ia), dt)
            b = random_array(rng, tuple(ext[c] for c in ib), dt)
            wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
            c = einsum_loops(list(ic), a.astype(wide), list(ia), b.astype(wide), list(ib)).astype(dt)
            key = f"{name}__{dt}"
            out[key + "__a"], out[key + "__b"], out[key + "__c"] = a, b, c
    path = os.path.join(HERE, "dangling_golden.npz")
    np.savez_compressed(path, **out)
    print(path, len(out) // 3, "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
