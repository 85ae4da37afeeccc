"""Generates tests/golden/binary_einsum_golden.npz.

Julia is not installed in the build image, so the reference itself cannot produce vectors. The golden
outputs are therefore produced by the plain-C loop nest (oracle/einsum_ref.c — the same explicit loop
the reference's own tests use as ground truth, test/integration/omeinsum.jl:225-236) on seeded inputs;
the NumPy/BLAS oracle and the CUDA path are both checked against them.

    python tests/golden/make_golden.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

from cases import PARITY_CASES, build_case  # noqa: E402
from oracle.build_oracle import einsum_loops  # noqa: E402

GOLDEN_CASES = ["matmul_small", "rank4_2sum", "mps_mpo_2b", "batch_peps", "two_batch", "outer", "inner",
                "size1_modes", "rank8_cfg5"]
DTYPES = ["float32", "float64", "complex64", "complex128"]


def main():
    out = {}
    for case in PARITY_CASES:
        if case[0] not in GOLDEN_CASES:
            continue
        for dt in DTYPES:
            a, ia, b, ib, ic = build_case(case, dt, seed=1234)
            # accumulate in double precision, then round once to the storage dtype
            wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
            c = einsum_loops(ic, a.astype(wide), ia, b.astype(wide), ib).astype(dt)
            key = f"{case[0]}__{dt}"
            out[key + "__a"] = a
            out[key + "__b"] = b
            out[key + "__c"] = c
    path = os.path.join(HERE, "binary_einsum_golden.npz")
    np.savez_compressed(path, **out)
    print(path, len(out) // 3, "cases", os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
