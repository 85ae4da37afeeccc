"""Shared contraction cases for the CPU and GPU test-suites.

REFERENCE_BATTERY transcribes the reference's own known-answer tests (all-ones inputs, exact `==`):
/root/reference/test/unit/operations/binary_einsum.jl:4-154, test/integration/omeinsum.jl:102-239,
test/integration/cuda.jl:8-182.  Each entry: (name, shape_a, inds_a, shape_b, inds_b, kwargs,
expected_inds, expected_shape, expected_value or None, needs_hyperindex).
"""
import numpy as np

REFERENCE_BATTERY = [
    # matmul — binary_einsum.jl:4-30
    ("matmul", (2, 3), "ij", (3, 4), "jk", {}, "ik", (2, 4), 3.0, False),
    ("matmul_out", (2, 3), "ij", (3, 4), "jk", {"out": "ik"}, "ik", (2, 4), 3.0, False),
    ("matmul_out_perm", (2, 3), "ij", (3, 4), "jk", {"out": "ki"}, "ki", (4, 2), 3.0, False),
    ("matmul_dims", (2, 3), "ij", (3, 4), "jk", {"dims": "j"}, "ik", (2, 4), 3.0, False),
    # inner product — :32-58
    ("inner", (3, 4), "ij", (4, 3), "ji", {}, "", (), 12.0, False),
    ("inner_out", (3, 4), "ij", (4, 3), "ji", {"out": ""}, "", (), 12.0, False),
    ("inner_dims", (3, 4), "ij", (4, 3), "ji", {"dims": "ij"}, "", (), 12.0, False),
    ("inner_dims_perm", (3, 4), "ij", (4, 3), "ji", {"dims": "ji"}, "", (), 12.0, False),
    # outer product — :60-94
    ("outer", (2, 3), "ij", (4, 5), "kl", {}, "ijkl", (2, 3, 4, 5), 1.0, False),
    ("outer_out", (2, 3), "ij", (4, 5), "kl", {"out": "ijkl"}, "ijkl", (2, 3, 4, 5), 1.0, False),
    ("outer_klij", (2, 3), "ij", (4, 5), "kl", {"out": "klij"}, "klij", (4, 5, 2, 3), 1.0, False),
    ("outer_lkji", (2, 3), "ij", (4, 5), "kl", {"out": "lkji"}, "lkji", (5, 4, 3, 2), 1.0, False),
    ("outer_likj", (2, 3), "ij", (4, 5), "kl", {"out": "likj"}, "likj", (5, 2, 4, 3), 1.0, False),
    ("outer_jikl", (2, 3), "ij", (4, 5), "kl", {"out": "jikl"}, "jikl", (3, 2, 4, 5), 1.0, False),
    # batch matmul — :122-133 (throws on BackendBase); omeinsum.jl:188-207, cuda.jl:126-144 (succeeds)
    ("batch_ikb", (2, 3, 6), "ijb", (3, 4, 6), "jkb", {"out": "ikb"}, "ikb", (2, 4, 6), 3.0, True),
    ("batch_kib", (2, 3, 6), "ijb", (3, 4, 6), "jkb", {"out": "kib"}, "kib", (4, 2, 6), 3.0, True),
    ("batch_bik", (2, 3, 6), "ijb", (3, 4, 6), "jkb", {"out": "bik"}, "bik", (6, 2, 4), 3.0, True),
    ("batch_dims", (2, 3, 6), "ijb", (3, 4, 6), "jkb", {"dims": "j"}, "ibk", (2, 6, 4), 3.0, True),
    # manual rank-3 — :135-154 ; all shared contracted → 12, `dims=[j]` → k is a hyperindex → 3
    ("manual_all", (2, 3, 4), "ijk", (4, 5, 3), "klj", {"dims": "jk"}, "il", (2, 5), 12.0, False),
    ("manual_hyper", (2, 3, 4), "ijk", (4, 5, 3), "klj", {"dims": "j"}, "ikl", (2, 4, 5), 3.0, True),
]


def _f(x):
    return x if x.ndim == 0 else np.asfortranarray(x)


def random_array(rng, shape, dtype):
    dtype = np.dtype(dtype)
    if dtype.kind == "c":
        real = np.float32 if dtype == np.complex64 else np.float64
        x = rng.uniform(-1, 1, size=shape).astype(real) + 1j * rng.uniform(-1, 1, size=shape).astype(real)
        return _f(np.asarray(x).astype(dtype))
    return _f(np.asarray(rng.uniform(-1, 1, size=shape)).astype(dtype))


def integer_array(rng, shape, dtype, lo=-3, hi=4):
    """Small-integer-valued data: every product and partial sum is exact in any order, so results can be
    compared with `==` (bit-exact index bookkeeping / output placement)."""
    dtype = np.dtype(dtype)
    x = rng.integers(lo, hi, size=shape).astype(np.float64)
    if dtype.kind == "c":
        x = x + 1j * rng.integers(lo, hi, size=shape).astype(np.float64)
    return _f(np.asarray(x).astype(dtype))


# (name, extents dict, inds_a, inds_b, inds_c) — random-data parity cases, sized for seconds on CPU
PARITY_CASES = [
    ("matmul_small", dict(i=5, j=7, k=3), "ij", "jk", "ik"),
    ("matmul_T", dict(i=33, j=17, k=65), "ji", "kj", "ki"),
    ("rank4_2sum", dict(a=6, b=5, c=7, d=4, e=3, f=8), "acbd", "dfce", "feab"),
    ("rank4_scrambled_cfg1", dict(i=12, j=10, k=9, l=11, m=8, n=13), "kilj", "nlmk", "mjni"),
    ("mps_mpo_2a", dict(a=24, w=4, b=20, s=2, c=16), "awb", "bsc", "awsc"),
    ("mps_mpo_2b", dict(a=24, w=4, s=2, c=16, t=2, v=4), "awsc", "wstv", "atvc"),
    ("mps_mpo_2c", dict(a=24, t=2, v=4, c=16, e=12), "atvc", "ate", "evc"),
    ("batch_peps", dict(l=6, k=3, b=4, m=10, q=5, r=7, z=3), "lkbmz", "mkqrz", "lbqrz"),
    ("batch_front", dict(i=9, j=8, k=7, z=5), "zij", "jzk", "kzi"),
    ("two_batch", dict(i=4, j=6, k=5, y=3, z=2), "iyjz", "zjky", "yikz"),
    ("outer", dict(i=6, j=5, k=4, l=3), "ij", "kl", "likj"),
    ("inner", dict(i=37, j=29), "ij", "ji", ""),
    ("vec_mat", dict(i=130, j=70), "i", "ij", "j"),
    ("mat_vec", dict(i=130, j=70), "ij", "j", "i"),
    ("scale_right", dict(i=6, j=5), "ij", "", "ji"),
    ("scale_left", dict(i=6, j=5), "", "ij", "ij"),
    ("size1_modes", dict(i=1, j=6, k=1, l=5, m=1), "ijkm", "klm", "lij"),
    ("rank6_cfg4", dict(a=4, b=3, c=5, d=4, e=3, f=5, g=2, h=3, p=4), "adbecf", "fgdhep", "pgachb"),
    ("rank8_cfg5", dict(a=2, b=3, c=2, d=3, e=2, f=3, g=2, h=3, p=2, q=3, r=2, s=3), "aebfcgdh", "hpgqfres", "srqpdcba"),
    ("skinny_n", dict(a=300, w=3, s=2, t=2, v=3), "aws", "wstv", "atv"),
    ("skinny_m", dict(a=300, w=3, s=2, t=2, v=3), "wstv", "aws", "tva"),
    ("big_k", dict(i=20, k=700, j=24), "ki", "kj", "ij"),
    ("tails", dict(i=131, j=67, k=19), "ik", "kj", "ij"),
]


def build_case(case, dtype, seed=0, integer=False):
    name, ext, ia, ib, ic = case
    rng = np.random.default_rng(seed)
    gen = integer_array if integer else random_array
    a = gen(rng, tuple(ext[c] for c in ia), dtype)
    b = gen(rng, tuple(ext[c] for c in ib), dtype)
    return a, list(ia), b, list(ib), list(ic)
