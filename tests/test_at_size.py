"""GPU parity AT BENCHMARK SIZE for every BASELINE.json config ("parity check runs on the full result
for cfgs 1-3, 4a, 5 and on a random 1/64 slab of C for 4b"). The CUDA path runs the full-size contraction through the
public API; the oracle is Muscle's BackendBase restated (oracle.binary_einsum_base: permutedims + BLAS gemm +
permutedims, src/Operations/binary_einsum.jl:76-96) in FP64 on the host, or the general (hyperindex) einsum for the
batched config. Tolerances are the north star's: rel. Frobenius <= 1e-12 (ComplexF64), <= 1e-5 (ComplexF32).
"""
import numpy as np
import pytest

import muscle_b200 as mb
from muscle_b200 import B200Array, Index, Tensor, _lib, binary_einsum
from cases import random_array
from oracle import binary_einsum_base, rel_frobenius

pytestmark = pytest.mark.gpu


def I(s):
    return [Index(c) for c in s]


def _dev_uniform(shape, dtype, seed):
    """uniform[-1,1) re/im generated on the device (seeded torch generator): (Tensor-ready B200Array, torch view with
    numpy's axis order reversed, i.e. t[..., i1, i0] is element (i0, i1, ...) of the column-major array)."""
    import torch
    g = torch.Generator(device="cuda:0")
    g.manual_seed(seed)
    n = int(np.prod(shape))
    real = torch.float64 if dtype == "complex128" else torch.float32
    flat = torch.rand(2 * n, dtype=real, device="cuda:0", generator=g) * 2 - 1
    arr = B200Array.from_torch(flat, shape, dtype)
    view = torch.view_as_complex(flat.view(n, 2)).view(*reversed(shape))
    return arr, view


def _to_numpy_colmajor(tview):
    """torch view in reversed axis order -> numpy array in the column-major axis order."""
    return np.asfortranarray(tview.contiguous().cpu().numpy().transpose())


def test_cfg1_rank4_dim64_full_vs_oracle():
    """configs[0]: two random ComplexF64 rank-4 tensors, dim 64, two summed indices, scrambled layout
    A[k,i,l,j] B[n,l,m,k] -> C[m,j,n,i]. Full result against the BackendBase restatement."""
    n = 64
    A = random_array(np.random.default_rng(1000), (n, n, n, n), "complex128")
    B = random_array(np.random.default_rng(1001), (n, n, n, n), "complex128")
    got = binary_einsum(Tensor(A, I("kilj")).to_device(), Tensor(B, I("nlmk")).to_device(), out=I("mjni")).to_host().data
    ref = binary_einsum_base(list("mjni"), A, list("kilj"), B, list("nlmk"))
    assert got.shape == ref.shape == (n, n, n, n)
    assert rel_frobenius(got, ref) <= 1e-12


def test_cfg4a_rank6_dim16_full_vs_oracle():
    """configs[3], literal reading: rank-6 ComplexF64, dim 16 per index, 3 summed (4096^3 GEMM-equivalent), interleaved
    labels. Full result against the BackendBase restatement."""
    n = 16
    ia, ib, ic = "adbecf", "fgdhei", "abcghi"
    A = random_array(np.random.default_rng(4000), (n,) * 6, "complex128")
    B = random_array(np.random.default_rng(4001), (n,) * 6, "complex128")
    got = binary_einsum(Tensor(A, I(ia)).to_device(), Tensor(B, I(ib)).to_device(), out=I(ic)).to_host().data
    ref = binary_einsum_base(list(ic), A, list(ia), B, list(ib))
    assert rel_frobenius(got, ref) <= 1e-12


def test_cfg4b_rank6_16k_random_slab_vs_oracle():
    """configs[3], the "~16k x 16k x 16k" reading (the bench headline): rank-6 ComplexF64, extents (32,32,16 | 32,32,16 |
    32,32,16) = 16384^3, 35.2 TFLOP on the device at full size; a random 1/64 slab of C (4 of 32 values of the free label
    `a` x 4 of 32 values of the free label `g`) against the oracle run on the matching operand slabs."""
    import torch
    ext = dict(a=32, b=32, c=16, d=32, e=32, f=16, g=32, h=32, i=16)
    ia, ib, ic = "adbecf", "fgdhei", "abcghi"
    A, vA = _dev_uniform([ext[c] for c in ia], "complex128", 4100)
    B, vB = _dev_uniform([ext[c] for c in ib], "complex128", 4101)
    Cc = binary_einsum(Tensor(A, I(ia)), Tensor(B, I(ib)), out=I(ic))
    assert Cc.shape == tuple(ext[c] for c in ic)
    rng = np.random.default_rng(4102)
    sa = np.sort(rng.choice(ext["a"], 4, replace=False))
    sg = np.sort(rng.choice(ext["g"], 4, replace=False))
    ta = torch.as_tensor(sa, device="cuda:0")
    tg = torch.as_tensor(sg, device="cuda:0")
    # torch views carry the axes reversed: A[a,d,b,e,c,f] -> vA[f,c,e,b,d,a]
    A_s = _to_numpy_colmajor(vA.index_select(5, ta))                    # a restricted
    B_s = _to_numpy_colmajor(vB.index_select(4, tg))                    # B[f,g,d,h,e,i] -> vB[i,e,h,d,g,f]: g is axis 4
    n = int(np.prod(Cc.shape))
    vC = torch.view_as_complex(Cc.data._owner[: 16 * n].view(torch.float64).view(n, 2)).view(*reversed(Cc.shape))
    C_s = _to_numpy_colmajor(vC.index_select(5, ta).index_select(2, tg))  # C[a,b,c,g,h,i] -> vC[i,h,g,c,b,a]
    del A, B, Cc, vA, vB, vC
    torch.cuda.empty_cache()
    ref = binary_einsum_base(list(ic), A_s, list(ia), B_s, list(ib))
    assert C_s.shape == ref.shape == (4, 32, 16, 4, 32, 16)
    assert rel_frobenius(C_s, ref) <= 1e-12


def test_cfg5_rank8_dim8_full_vs_fp64_oracle():
    """configs[4] on one GPU (the summed-index slice over ranks is tests/test_multi_gpu.py): rank-8 ComplexF32, dim 8,
    4 summed (4096^3), labels interleaved, output reversed. Full result against the FP64 oracle, <= 1e-5."""
    n = 8
    ia, ib, ic = "aebfcgdh", "hpgqfres", "srqpdcba"
    A = random_array(np.random.default_rng(5000), (n,) * 8, "complex64")
    B = random_array(np.random.default_rng(5100), (n,) * 8, "complex64")
    h = _lib.Handle.get()
    h.reset_stats()
    got = binary_einsum(Tensor(A, I(ia)).to_device(), Tensor(B, I(ib)).to_device(), out=I(ic)).to_host().data
    assert h.stats()["launches_tcgen05"] == 1          # the tensor-core path is the one that ran
    ref = binary_einsum_base(list(ic), A.astype(np.complex128), list(ia), B.astype(np.complex128), list(ib))
    assert rel_frobenius(got.astype(np.complex128), ref) <= 1e-5


def test_cfg3_peps_full_size_every_batch_vs_fp64_oracle():
    """configs[2] at full size (ComplexF32, D = 8, chi = 256, beta = 8: 2048^3 x 8): EVERY batch slice against the FP64
    oracle (BackendBase rejects hyperindices, so the oracle contracts slice by slice - the loop the reference's own
    hyperindex test uses, test/integration/cuda.jl:169-180)."""
    chi, D, beta = 256, 8, 8
    A = random_array(np.random.default_rng(3000), (chi, D, D, chi, beta), "complex64")
    B = random_array(np.random.default_rng(3001), (chi, D, D, chi, beta), "complex64")
    got = binary_einsum(Tensor(A, I("lkbmz")).to_device(), Tensor(B, I("mkqrz")).to_device(), out=I("lbqrz")).to_host().data
    num = den = 0.0
    for z in range(beta):
        ref = binary_einsum_base(list("lbqr"), A[..., z].astype(np.complex128), list("lkbm"),
                                 B[..., z].astype(np.complex128), list("mkqr"))
        num += float(np.linalg.norm((got[..., z] - ref).ravel()) ** 2)
        den += float(np.linalg.norm(ref.ravel()) ** 2)
    assert np.sqrt(num / den) <= 1e-5


# ---- strict-FP32 compute type and the exactness limit of the split scheme -------------------------------
def test_compute_type_fp32_is_exact_on_integers_above_2_pow_11():
    """Integers above 2^11 do not fit TF32's significand: the default tensor-core split scheme drops lo*lo and rounds the
    cross terms to bf16, so products such as 4097 * 2049 (exact in FP32) can come out off by one. COMPUTE_FP32 routes the
    same contraction to FP32 FMAs and reproduces the exact integers; the default stays within the 1e-5 contract."""
    m, n, k = 512, 512, 512
    rng = np.random.default_rng(77)
    a = np.zeros((k, m), np.float32)
    b = np.zeros((k, n), np.float32)
    # one non-zero per output dot product: the result is a single product, exact in FP32 (< 2^24)
    a[np.arange(m) % k, np.arange(m)] = rng.integers(2049, 4096, m).astype(np.float32)
    b[:, :] = 0
    b[np.arange(k), :] = rng.integers(2049, 4096, (k, 1)).astype(np.float32)
    ref = a.astype(np.float64).T @ b.astype(np.float64)
    assert np.abs(ref).max() < 2 ** 24
    h = _lib.Handle.get()
    ta, tb = Tensor(a, I("ki")).to_device(), Tensor(b, I("kj")).to_device()
    try:
        h.set_compute_type(_lib.COMPUTE_FP32)
        assert h.compute_type() == _lib.COMPUTE_FP32
        h.reset_stats()
        exact = binary_einsum(ta, tb, out=I("ij")).to_host().data
        s = h.stats()
        assert s["launches_tcgen05"] == 0 and s["launches_simt_f32"] == 1
        assert np.array_equal(exact.astype(np.float64), ref)
        h.set_compute_type(_lib.COMPUTE_DEFAULT)
        h.reset_stats()
        split = binary_einsum(ta, tb, out=I("ij")).to_host().data
        assert h.stats()["launches_tcgen05"] == 1
        assert rel_frobenius(split.astype(np.float64), ref) <= 1e-5
        h.set_compute_type(_lib.COMPUTE_3XTF32)
        three = binary_einsum(ta, tb, out=I("ij")).to_host().data
        assert rel_frobenius(three.astype(np.float64), ref) <= 1e-5
    finally:
        h.set_compute_type(_lib.COMPUTE_DEFAULT)
    with pytest.raises(mb.ArgumentError):
        h.set_compute_type(7)


# ---- dangling labels (one operand only, absent from C): summed, cuTENSOR / OMEinsum semantics ---------------------------
DANGLING = [
    # name, extents, inds_a, inds_b, inds_c                 (x, y: dangling)
    ("fold_a", dict(i=5, j=7, k=3, x=4), "ixj", "jk", "ik"),
    ("fold_b", dict(i=5, j=7, k=3, y=6), "ij", "yjk", "ki"),
    ("fold_both", dict(i=4, j=5, k=3, x=2, y=3), "xij", "jky", "ik"),
    ("fold_unit", dict(i=4, j=5, k=3, x=1), "ijx", "jk", "ik"),
    ("fold_to_scalar", dict(i=6, x=5), "ix", "i", ""),
    ("fold_batch", dict(i=4, j=5, k=3, z=2, x=3), "izxj", "jkz", "kiz"),
    ("prereduce_a", dict(i=96, j=80, k=72, x=24), "ixj", "jk", "ik"),
    ("prereduce_b_slowest", dict(i=130, j=70, k=90, y=9), "ji", "jky", "ik"),
    ("prereduce_both", dict(i=64, j=64, k=64, x=8, y=6), "xij", "yjk", "ki"),
    ("prereduce_tc", dict(i=512, j=256, k=384, x=4), "jxi", "jk", "ik"),
]


@pytest.mark.parametrize("device", [False, True], ids=["host_entry", "device_entry"])
@pytest.mark.parametrize("dt", ["float32", "float64", "complex64", "complex128"])
@pytest.mark.parametrize("case", DANGLING, ids=[c[0] for c in DANGLING])
def test_dangling_labels_are_summed(case, dt, device):
    from oracle import binary_einsum_general
    name, ext, ia, ib, ic = case
    rng = np.random.default_rng(11)
    a = random_array(rng, tuple(ext[c] for c in ia), dt)
    b = random_array(rng, tuple(ext[c] for c in ib), dt)
    ta, tb = Tensor(a, I(ia)), Tensor(b, I(ib))
    if device:
        ta, tb = ta.to_device(), tb.to_device()
    got = binary_einsum(mb.BackendB200(), I(ic), ta, tb).to_host().data
    wide = np.complex128 if "complex" in dt else np.float64
    ref = binary_einsum_general(list(ic), a.astype(wide), list(ia), b.astype(wide), list(ib))
    tol = 1e-12 if dt in ("float64", "complex128") else 1e-5
    assert got.shape == ref.shape
    assert rel_frobenius(got.astype(wide), ref) <= tol


def test_dangling_labels_integer_exact_and_mixed_eltypes():
    from oracle import binary_einsum_general
    rng = np.random.default_rng(5)
    a = rng.integers(-3, 4, (6, 4, 5)).astype(np.float64)
    b = (rng.integers(-3, 4, (5, 7, 3)) + 1j * rng.integers(-3, 4, (5, 7, 3))).astype(np.complex128)
    got = binary_einsum(mb.BackendB200(), I("ki"), Tensor(a, I("ixj")).to_device(), Tensor(b, I("jky")).to_device())
    assert got.dtype == np.complex128
    assert np.array_equal(got.to_host().data, binary_einsum_general(list("ki"), a, list("ixj"), b, list("jky")))


def test_operands_on_one_device_only():
    """Operands on different GPUs are an ArgumentError, never a launch with foreign pointers."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    a = Tensor(np.ones((4, 4)), I("ij")).to_device(0)
    b = Tensor(np.ones((4, 4)), I("jk")).to_device(1)
    with pytest.raises(mb.ArgumentError):
        binary_einsum(a, b)


# ---- cross-GPU split-K (fused contraction + all-reduce) with the ranks EMULATED on one GPU ------------------------------------
ALLREDUCE_EMU = [
    # dt, (M, N, K), nranks, out permuted?            pair kernel: M >= 256
    ("complex64", (512, 384, 1024), 4, False),
    ("complex64", (1000, 520, 2048), 8, True),        # ragged tiles, transposed output (non-contiguous C rows)
    ("complex64", (128, 256, 512), 2, False),         # single-CTA kernel (M < 256)
    ("float32", (768, 640, 1024), 4, False),
    ("float32", (520, 300, 2064), 3, True),
]


@pytest.mark.parametrize("dt,dims,nranks,permuted", ALLREDUCE_EMU)
def test_fused_allreduce_emulated_ranks(dt, dims, nranks, permuted):
    """mb200_binary_einsum_allreduce with the phases split so that one GPU can play every rank: all ranks contract their
    K-slice (partial units -> own workspace, unit flags -> owners), then every owner's reducer adds the partials of its units
    and stores them into the C of EVERY rank, then the done-flag waits. Every rank's C must equal the unsliced contraction
    (what `treereduce(AddComputeOp)` / all_reduce(SUM) of the partials gives) and all ranks must hold identical bits."""
    import ctypes as C
    Mx, Nx, Kx = dims
    rng = np.random.default_rng(23)
    a = random_array(rng, (Kx, Mx), dt)          # [k, i]
    b = random_array(rng, (Kx, Nx), dt)          # [k, j]
    wide = np.complex128 if dt == "complex64" else np.float64
    ref = a.astype(wide).T @ b.astype(wide)       # C[i, j]
    modes_c = [2, 1] if permuted else [1, 2]
    if permuted:
        ref = ref.T
    h = _lib.Handle.get()
    h.set_path(mb.PATH_TCGEN05_TF32)
    L = mb.lib()
    en = _lib.dtype_enum(dt)
    kc = Kx // nranks
    assert kc * nranks == Kx and kc % 8 == 0
    try:
        ws_b, fl_b = C.c_size_t(), C.c_size_t()
        _lib.check(L.mb200_allreduce_workspace(h.ptr, en, 2, _lib.i32(modes_c), en, 2, _lib.i32([0, 1]), _lib.i64((kc, Mx)),
                                               en, 2, _lib.i32([0, 2]), _lib.i64((kc, Nx)), nranks, C.byref(ws_b), C.byref(fl_b)))
        ws = [B200Array((ws_b.value,), np.float32) for _ in range(nranks)]          # oversized on purpose (bytes as floats)
        fl = [B200Array((fl_b.value // 4 + 1,), np.float32) for _ in range(nranks)]
        cs = [B200Array(ref.shape, dt) for _ in range(nranks)]
        for r in range(nranks):
            _lib.check(L.mb200_memset(h.ptr, C.c_void_p(fl[r].ptr), 0, fl[r].nbytes))
            _lib.check(L.mb200_memset(h.ptr, C.c_void_p(cs[r].ptr), 0xFF, cs[r].nbytes))   # NaN-fill: every element must be written
            _lib.check(L.mb200_memset(h.ptr, C.c_void_p(ws[r].ptr), 0xFF, ws[r].nbytes))
        slices = [(B200Array.from_host(a[r * kc:(r + 1) * kc, :]), B200Array.from_host(b[r * kc:(r + 1) * kc, :])) for r in range(nranks)]
        for epoch in (1, 2):                       # twice: flags carry epochs, nothing is reset between calls
            for phase in (_lib.DIST_CONTRACT, _lib.DIST_REDUCE, _lib.DIST_WAIT):
                for r in range(nranks):
                    cm = _lib.Comm()
                    cm.nranks, cm.rank, cm.epoch = nranks, r, epoch
                    for q in range(nranks):
                        cm.ws[q], cm.c[q], cm.flags[q] = ws[q].ptr, cs[q].ptr, fl[q].ptr
                    cm.ws_bytes, cm.flag_bytes = ws_b.value, fl_b.value
                    da, db = slices[r]
                    _lib.check(L.mb200_binary_einsum_allreduce(
                        h.ptr, en, 2, _lib.i32(modes_c),
                        C.c_void_p(da.ptr), en, 2, _lib.i32([0, 1]), _lib.i64(da.shape), None,
                        C.c_void_p(db.ptr), en, 2, _lib.i32([0, 2]), _lib.i64(db.shape), None, C.byref(cm), phase))
            got = [c.to_host() for c in cs]
            assert rel_frobenius(got[0].astype(wide), ref) <= 1e-5
            for r in range(1, nranks):
                assert np.array_equal(got[r], got[0])
            if epoch == 1:
                for r in range(nranks):
                    _lib.check(L.mb200_memset(h.ptr, C.c_void_p(cs[r].ptr), 0xFF, cs[r].nbytes))
        # the production call (all three phases at once, reducer on the side stream) with ONE rank: plain contraction semantics
        cm = _lib.Comm()
        cm.nranks, cm.rank, cm.epoch = 1, 0, 3
        cm.ws[0], cm.c[0], cm.flags[0] = ws[0].ptr, cs[0].ptr, fl[0].ptr
        cm.ws_bytes, cm.flag_bytes = ws_b.value, fl_b.value
        _lib.check(L.mb200_memset(h.ptr, C.c_void_p(fl[0].ptr), 0, fl[0].nbytes))
        _lib.check(L.mb200_memset(h.ptr, C.c_void_p(cs[0].ptr), 0xFF, cs[0].nbytes))
        da, db = B200Array.from_host(a), B200Array.from_host(b)
        _lib.check(L.mb200_binary_einsum_allreduce(
            h.ptr, en, 2, _lib.i32(modes_c), C.c_void_p(da.ptr), en, 2, _lib.i32([0, 1]), _lib.i64(da.shape), None,
            C.c_void_p(db.ptr), en, 2, _lib.i32([0, 2]), _lib.i64(db.shape), None, C.byref(cm), 7))
        assert rel_frobenius(cs[0].to_host().astype(wide), ref) <= 1e-5
    finally:
        h.set_path(mb.PATH_AUTO)


def test_fused_allreduce_rejects_other_paths():
    import ctypes as C
    h = _lib.Handle.get()
    ws_b, fl_b = C.c_size_t(), C.c_size_t()
    with pytest.raises(mb.ArgumentError):      # ComplexF64 runs on DMMA, not on the tcgen05 path
        _lib.check(mb.lib().mb200_allreduce_workspace(h.ptr, _lib.C128, 2, _lib.i32([1, 2]), _lib.C128, 2, _lib.i32([0, 1]),
                                                      _lib.i64((512, 512)), _lib.C128, 2, _lib.i32([0, 2]), _lib.i64((512, 512)),
                                                      4, C.byref(ws_b), C.byref(fl_b)))


def test_dangling_golden_vectors():
    """The CUDA path against tests/golden/dangling_golden.npz (outputs of the plain-C loop nest): both routes - the sum folded into the
    direct kernel, and the unary_einsum pre-reduce for the case above 2^20 MACs."""
    import importlib.util
    import os
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("mgd", os.path.join(here, "golden", "make_golden_dangling.py"))
    mgd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgd)
    z = np.load(os.path.join(here, "golden", "dangling_golden.npz"))
    h = _lib.Handle.get()
    for name, ext, ia, ib, ic in mgd.CASES:
        for dt in mgd.DTYPES:
            a, b, c = z[f"{name}__{dt}__a"], z[f"{name}__{dt}__b"], z[f"{name}__{dt}__c"]
            h.reset_stats()
            got = binary_einsum(mb.BackendB200(), I(ic), Tensor(a, I(ia)).to_device(), Tensor(b, I(ib)).to_device()).to_host().data
            st = h.stats()
            assert (st["launches_unary"] > 0) == (name == "prereduce_a"), (name, st)     # which route ran
            wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
            tol = 1e-12 if dt in ("float64", "complex128") else 1e-5
            assert got.shape == c.shape and rel_frobenius(got.astype(wide), c.astype(wide)) <= tol, (name, dt)
