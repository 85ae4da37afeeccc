"""GPU parity tests: the CUDA path (through the C ABI) against the oracle on the same seeded inputs,
the reference's known-answer battery replayed on BackendB200, the committed golden vectors, and
size-independent properties at the BASELINE.json sizes.

Tolerances (north star): rel. Frobenius <= 1e-12 for Float64/ComplexF64, <= 1e-5 for Float32/ComplexF32;
index bookkeeping / output placement bit-exact (integer-valued inputs, compared with ==).
"""
import ctypes as C
import os

import numpy as np
import pytest

import muscle_b200 as mb
from muscle_b200 import B200Array, BackendB200, Index, Tensor, _lib, binary_einsum, binary_einsum_, with_backend
from cases import PARITY_CASES, REFERENCE_BATTERY, build_case, integer_array, random_array
from oracle import binary_einsum_general, rel_frobenius

pytestmark = pytest.mark.gpu

TOL = {"float32": 1e-5, "complex64": 1e-5, "float64": 1e-12, "complex128": 1e-12}
DTYPES = ["float32", "float64", "complex64", "complex128"]
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "binary_einsum_golden.npz")


def I(s):
    return [Index(c) for c in s]


@pytest.fixture(autouse=True)
def _auto_path():
    yield
    _lib.Handle.get().set_path(mb.PATH_AUTO)


def contract(a, ia, b, ib, ic, device=True, path=mb.PATH_AUTO):
    _lib.Handle.get().set_path(path)
    ta, tb = Tensor(a, I(ia)), Tensor(b, I(ib))
    if device:
        ta, tb = ta.to_device(), tb.to_device()
    c = binary_einsum(BackendB200(), I(ic), ta, tb)
    assert c.inds == I(ic)
    assert c.on_device == device
    return c.to_host().data


# ---- the reference's own battery on the new backend --------------------------------------------
@pytest.mark.parametrize("device", [False, True], ids=["host_entry", "device_entry"])
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", REFERENCE_BATTERY, ids=[c[0] for c in REFERENCE_BATTERY])
def test_reference_battery(case, dt, device):
    """test/unit/operations/binary_einsum.jl:4-154 re-run per backend, the way
    test/integration/{strided,omeinsum,cuda}.jl do; hyperindex cases succeed as on cuTENSOR/OMEinsum."""
    name, sa, ia, sb, ib, kw, exp_inds, exp_shape, exp_val, hyper = case
    A, B = Tensor(np.ones(sa, dt), I(ia)), Tensor(np.ones(sb, dt), I(ib))
    if device:
        A, B = A.to_device(), B.to_device()
    kwargs = {k: I(v) for k, v in kw.items()}
    Cc = with_backend(lambda: binary_einsum(A, B, **kwargs), BackendB200())
    assert Cc.inds == I(exp_inds)
    assert Cc.shape == exp_shape
    got = Cc.to_host().data
    assert got.dtype == np.dtype(dt)
    assert np.array_equal(got, np.full(exp_shape, exp_val, dt))


def test_reference_scale_and_mixed_eltypes():
    """scale — binary_einsum.jl:96-119; mixed ComplexF64 × Float64 — omeinsum.jl:160-186, cuda.jl:100-124."""
    for dt in (np.float64, np.complex128, np.float32, np.complex64):
        A = Tensor(np.ones((2, 3), dt), I("ij"))
        alpha = Tensor(np.array(2.0))
        for dev in (False, True):
            a, al = (A.to_device(), alpha.to_device()) if dev else (A, alpha)
            for x, y in ((a, al), (al, a)):
                Cc = with_backend(lambda: binary_einsum(x, y), BackendB200())
                assert Cc.inds == I("ij")
                exp_dt = np.result_type(dt, np.float64)
                got = Cc.to_host().data
                assert got.dtype == exp_dt and np.array_equal(got, 2.0 * np.ones((2, 3), exp_dt))
                Cc = with_backend(lambda: binary_einsum(x, y, out=I("ji")), BackendB200())
                assert Cc.shape == (3, 2) and np.array_equal(Cc.to_host().data, 2.0 * np.ones((3, 2), exp_dt))


def test_auto_dispatch_on_device_arrays():
    """Domain(B200Array) → BackendB200 without with_backend (pattern ext/MuscleCUDAExt.jl:7, binary_einsum.jl:21);
    host/device "hybrid" operands (test/integration/reactant.jl:36-39)."""
    A = Tensor(np.ones((2, 3)), I("ij"))
    B = Tensor(np.ones((3, 4)), I("jk"))
    for x, y in ((A.to_device(), B.to_device()), (A.to_device(), B), (A, B.to_device())):
        Cc = binary_einsum(x, y)
        assert Cc.on_device and np.array_equal(Cc.to_host().data, 3 * np.ones((2, 4)))


def test_inplace():
    """binary_einsum!(c, a, b) writes parent(c) in inds(c) order and returns c (binary_einsum.jl:57-70)."""
    rng = np.random.default_rng(5)
    a, b = random_array(rng, (6, 7, 5), "complex128"), random_array(rng, (5, 4, 7), "complex128")
    ref = binary_einsum_general(list("li"), a, list("ijk"), b, list("klj"))
    Cd = Tensor(B200Array((4, 6), "complex128"), I("li"))
    out = binary_einsum_(Cd, Tensor(a, I("ijk")).to_device(), Tensor(b, I("klj")).to_device())
    assert out is Cd and rel_frobenius(Cd.to_host().data, ref) <= 1e-12
    Ch = Tensor(np.zeros((4, 6), np.complex128, order="F"), I("li"))
    with_backend(lambda: binary_einsum_(Ch, Tensor(a, I("ijk")), Tensor(b, I("klj"))), BackendB200())
    assert rel_frobenius(Ch.data, ref) <= 1e-12


# ---- random-data parity against the oracle --------------------------------------------------------
@pytest.mark.parametrize("path", ["auto", "direct", "tiled"])
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", PARITY_CASES, ids=[c[0] for c in PARITY_CASES])
def test_parity_random(case, dt, path):
    a, ia, b, ib, ic = build_case(case, dt, seed=11)
    ref = binary_einsum_general(ic, a, ia, b, ib)
    p = {"auto": mb.PATH_AUTO, "direct": mb.PATH_DIRECT,
         "tiled": mb.PATH_GETT_F64 if np.dtype(dt).itemsize >= 8 and dt != "complex64" else mb.PATH_SIMT_F32}[path]
    got = contract(a, ia, b, ib, ic, device=True, path=p)
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert rel_frobenius(got, ref) <= TOL[dt], (case[0], dt, path)


@pytest.mark.parametrize("path", ["direct", "tiled"])
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("case", PARITY_CASES, ids=[c[0] for c in PARITY_CASES])
def test_bookkeeping_bit_exact(case, dt, path):
    """Integer-valued inputs: every product and sum is exact, so any misplaced element, wrong stride or
    wrong classification shows up as inequality. Bit-exact (==) against the oracle."""
    a, ia, b, ib, ic = build_case(case, dt, seed=23, integer=True)
    ref = binary_einsum_general(ic, a, ia, b, ib)
    p = {"direct": mb.PATH_DIRECT,
         "tiled": mb.PATH_GETT_F64 if np.dtype(dt).itemsize >= 8 and dt != "complex64" else mb.PATH_SIMT_F32}[path]
    got = contract(a, ia, b, ib, ic, device=True, path=p)
    assert np.array_equal(got, ref), (case[0], dt, path)


def test_golden_vectors():
    g = np.load(GOLDEN)
    by_name = {c[0]: c for c in PARITY_CASES}
    keys = sorted({k.rsplit("__", 1)[0] for k in g.files})
    for key in keys:
        name, dt = key.split("__")
        _, _, ia, ib, ic = by_name[name]
        for device in (False, True):
            got = contract(g[key + "__a"], ia, g[key + "__b"], ib, ic, device=device)
            assert got.shape == g[key + "__c"].shape
            assert rel_frobenius(got, g[key + "__c"]) <= TOL[dt], key


@pytest.mark.parametrize("dt", DTYPES)
def test_multi_tile_shapes(dt):
    """Several CTA tiles, ragged edges, grouped rasterisation, K not a multiple of the k-block."""
    rng = np.random.default_rng(2)
    for (m, n, k, l) in [(1000, 520, 300, 1), (257, 129, 65, 3), (4100, 70, 33, 1), (70, 4100, 33, 1),
                         (130, 18, 1030, 2), (18, 130, 1030, 2)]:
        a = random_array(rng, (k, m, l), dt)
        b = random_array(rng, (n, l, k), dt)
        ref = np.einsum("kml,nlk->nml", a, b, optimize=True)
        got = contract(a, "kml", b, "nlk", "nml")
        assert rel_frobenius(got, ref) <= TOL[dt], (m, n, k, l, dt)
        got = contract(a, "kml", b, "nlk", "mln")
        assert rel_frobenius(got, np.transpose(ref, (1, 2, 0))) <= TOL[dt], (m, n, k, l, dt)


def test_mixed_eltypes_promote():
    rng = np.random.default_rng(9)
    for da, db in [("float64", "complex128"), ("complex64", "float32"), ("float32", "float64"),
                   ("complex64", "float64"), ("complex128", "complex64")]:
        a, b = random_array(rng, (40, 30), da), random_array(rng, (30, 50), db)
        ref = binary_einsum_general(list("ik"), a, list("ij"), b, list("jk"))
        for path in (mb.PATH_DIRECT, mb.PATH_AUTO):
            got = contract(a, "ij", b, "jk", "ik", path=path)
            assert got.dtype == ref.dtype
            assert rel_frobenius(got, ref) <= (1e-5 if ref.dtype.itemsize <= 8 and ref.dtype != np.float64 else 1e-12)


def test_strided_operands_through_c_abi():
    """A slab of a larger array (the free-index shard case: strides passed explicitly, no copy)."""
    rng = np.random.default_rng(4)
    a_full = random_array(rng, (12, 9, 10), "complex128")     # [i, k, s]
    b = random_array(rng, (9, 7), "complex128")               # [k, j]
    dA, dB = B200Array.from_host(a_full), B200Array.from_host(b)
    dC = B200Array((12, 7, 4), "complex128")                  # [i, j, s] for s in 3:7
    h = _lib.Handle.get()
    s0 = 3
    ptrA = dA.ptr + s0 * 12 * 9 * 16
    _lib.check(mb.lib().mb200_binary_einsum(
        h.ptr, C.c_void_p(dC.ptr), _lib.C128, 3, _lib.i32([0, 2, 3]), None,
        C.c_void_p(ptrA), _lib.C128, 3, _lib.i32([0, 1, 3]), _lib.i64([12, 9, 4]), _lib.i64([1, 12, 108]),
        C.c_void_p(dB.ptr), _lib.C128, 2, _lib.i32([1, 2]), _lib.i64([9, 7]), None))
    ref = np.einsum("iks,kj->ijs", a_full[:, :, 3:7], b)
    assert rel_frobenius(dC.to_host(), ref) <= 1e-12
    # strided C: write every other column of a wider buffer
    dC2 = B200Array((12, 14), "complex128")
    _lib.check(mb.lib().mb200_memset(h.ptr, C.c_void_p(dC2.ptr), 0, dC2.nbytes))
    a2 = random_array(rng, (12, 9), "complex128")
    dA2 = B200Array.from_host(a2)
    _lib.check(mb.lib().mb200_binary_einsum(
        h.ptr, C.c_void_p(dC2.ptr), _lib.C128, 2, _lib.i32([0, 2]), _lib.i64([1, 24]),
        C.c_void_p(dA2.ptr), _lib.C128, 2, _lib.i32([0, 1]), _lib.i64([12, 9]), None,
        C.c_void_p(dB.ptr), _lib.C128, 2, _lib.i32([1, 2]), _lib.i64([9, 7]), None))
    got = dC2.to_host()
    assert rel_frobenius(got[:, 0::2], a2 @ b) <= 1e-12 and not got[:, 1::2].any()


def test_empty_and_degenerate():
    a = np.ones((3, 0), np.float64, order="F")
    b = np.ones((0, 4), np.float64, order="F")
    got = contract(a, "ij", b, "jk", "ik")          # empty contraction → zeros
    assert got.shape == (3, 4) and not got.any()
    got = contract(np.ones((0, 3)), "ij", np.ones((3, 4)), "jk", "ik")
    assert got.shape == (0, 4)
    x = np.array(3.0 + 1j)
    y = np.array(2.0 - 1j)
    got = contract(x, "", y, "", "")                # scalar × scalar
    assert got.shape == () and got == x * y


def test_error_behaviour_on_device():
    A = Tensor(np.ones((2, 3)), I("ij")).to_device()
    B = Tensor(np.ones((3, 4)), I("jk")).to_device()
    got = binary_einsum(BackendB200(), I("i"), A, B)      # free label missing from C: summed (cuTENSOR semantics)
    assert np.array_equal(got.to_host().data, 12 * np.ones(2))
    with pytest.raises(mb.ArgumentError):           # label of C in neither operand
        binary_einsum(BackendB200(), I("iz"), A, B)
    with pytest.raises(mb.DimensionMismatch):
        binary_einsum(A, Tensor(np.ones((5, 4)), I("jk")).to_device())


# ---- K1 permute kernel: bit-exact ----------------------------------------------------------------------
PERMUTE_CASES = [
    ((64, 48), (1, 0)), ((33, 17, 9), (2, 0, 1)), ((33, 17, 9), (1, 2, 0)), ((33, 17, 9), (0, 2, 1)),
    ((2, 40, 3, 24), (3, 1, 0, 2)), ((2, 40, 3, 24), (1, 0, 3, 2)), ((16, 16, 16, 16), (1, 3, 0, 2)),
    ((16, 16, 16, 16), (3, 2, 1, 0)), ((5, 1, 7, 1, 3), (4, 3, 2, 1, 0)), ((300, 2), (1, 0)), ((2, 300), (1, 0)),
    ((2, 2, 2, 2, 2, 2, 2, 2, 2, 2), (9, 0, 8, 1, 7, 2, 6, 3, 5, 4)), ((12, 10, 8), (0, 1, 2)), ((1000,), (0,)),
    ((7, 130, 5), (0, 2, 1)), ((200, 3, 50), (0, 2, 1)),
    # register-tile transposition (extents % 4, % 2), ragged tiles, long y tables, widened unit-stride runs
    ((260, 132), (1, 0)), ((2, 1030, 6), (1, 0, 2)), ((8, 8, 8, 8, 8), (4, 0, 3, 1, 2)), ((4, 6, 10), (2, 1, 0)),
    ((6, 10, 14), (1, 0, 2)), ((4, 33, 5), (0, 2, 1)), ((8, 10, 12), (0, 2, 1)), ((6, 10, 12), (0, 2, 1)),
    ((2, 2048, 4), (1, 2, 0)), ((1028, 4, 8), (2, 1, 0)),
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,perm", PERMUTE_CASES)
def test_permute_bit_exact(shape, perm, dt):
    """`permutedims(t, perm)` (src/Tensor.jl:302-308 → Base.permutedims) is pure data movement: ==."""
    rng = np.random.default_rng(1)
    x = random_array(rng, shape, dt)
    t = Tensor(x, [Index(i) for i in range(len(shape))]).to_device()
    got = t.permutedims(list(perm))
    assert got.inds == [Index(p) for p in perm]
    assert np.array_equal(got.to_host().data, np.transpose(x, perm))


TMA_PERMUTE_CASES = [
    ((64, 64, 64, 64), (1, 3, 0, 2)),        # config-1 A pack: k,i,l,j -> i,j,k,l
    ((4096, 512), (1, 0)), ((100, 96), (1, 0)), ((33, 70, 40), (2, 1, 0)), ((48, 3, 80, 2), (2, 1, 0, 3)),
    ((256, 8, 8, 256, 2), (3, 1, 0, 2, 4)),  # config-3 A matricise (5 modes)
    ((40, 36, 5, 3, 2), (1, 0, 4, 3, 2)), ((130, 34), (1, 0)), ((32, 32), (1, 0)), ((64, 2, 64), (2, 1, 0)),
]


@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,perm", TMA_PERMUTE_CASES)
def test_permute_tma_variant_bit_exact(shape, perm, dt):
    """The TMA-staged transposition kernel (cp.async.bulk.tensor load -> shared-memory transposition -> bulk tensor store),
    forced through MB200_PERMUTE_TMA: bit-exact against numpy, ragged edges (zero-filled loads, clipped stores), 4 / 8 / 16-byte
    elements, up to 5 modes; ineligible layouts fall back silently and must still be exact."""
    rng = np.random.default_rng(3)
    a = random_array(rng, shape, dt)
    t = Tensor(a, [Index(i) for i in range(len(shape))]).to_device()
    got = t.permutedims(list(perm), flags=2).to_host().data
    assert np.array_equal(got, np.transpose(a, perm))
    ref = t.permutedims(list(perm), flags=4).to_host().data
    assert np.array_equal(ref, got)


def test_permute_unaligned_pointers_skip_the_wide_paths():
    """16-byte vector paths (element widening, register-tile transposition) need 16-byte aligned buffers: a Float32
    view that starts 4 bytes into an allocation must fall back to the element-wise kernels and still be exact."""
    rng = np.random.default_rng(2)
    for shape, perm in (((64, 48), (1, 0)), ((8, 12, 10), (0, 2, 1)), ((16, 16, 16), (2, 1, 0))):
        n = int(np.prod(shape))
        x = np.asfortranarray(rng.integers(-9, 9, size=shape).astype(np.float32))
        src = B200Array((n + 4,), np.float32)
        dst = B200Array((n + 4,), np.float32)
        flat = np.zeros(n + 4, np.float32)
        flat[1:n + 1] = x.ravel(order="F")
        src.copy_from_host(flat)
        h = _lib.Handle.get()
        _lib.check(mb.lib().mb200_permute(h.ptr, C.c_void_p(dst.ptr + 4), C.c_void_p(src.ptr + 4), _lib.F32, len(shape),
                                          _lib.i64(shape), _lib.i32(perm), 0))
        got = dst.to_host()[1:n + 1].reshape([shape[p] for p in perm], order="F")
        assert np.array_equal(got, np.transpose(x, perm))


def test_permute_planar_split():
    """complex interleaved → planar (all re, then all im) in the same pass."""
    rng = np.random.default_rng(1)
    for dt, real in (("complex128", np.float64), ("complex64", np.float32)):
        x = random_array(rng, (24, 10, 18), dt)
        src = B200Array.from_host(x)
        dst = B200Array((2, 18, 24, 10), real)   # planes slowest in memory → shape (18,24,10,2) col-major
        dst = B200Array((18, 24, 10, 2), real)
        h = _lib.Handle.get()
        _lib.check(mb.lib().mb200_permute(h.ptr, C.c_void_p(dst.ptr), C.c_void_p(src.ptr), _lib.dtype_enum(dt), 3,
                                          _lib.i64(x.shape), _lib.i32([2, 0, 1]), 1))
        got = dst.to_host()
        ref = np.transpose(x, (2, 0, 1))
        assert np.array_equal(got[..., 0], ref.real) and np.array_equal(got[..., 1], ref.imag)


# ---- BASELINE.json configs at full size ----------------------------------------------------------------
def _chain_cfg2(E, A, W, Ab):
    T = binary_einsum(E, A, out=I("awsc"))
    T2 = binary_einsum(T, W, out=I("atvc"))
    return binary_einsum(T2, Ab, out=I("evc"))


def test_cfg2_mps_mpo_full_size_vs_oracle():
    """configs[1]: MPS–MPO transfer contraction ComplexF64, χ=1024, d=2, MPO bond 8 — the bench workload.
    Full-size result against the oracle (CPU BLAS), all three steps."""
    chi, d, w = 1024, 2, 8
    rng = np.random.default_rng(2000)
    E = random_array(rng, (chi, w, chi), "complex128")
    A = random_array(np.random.default_rng(2001), (chi, d, chi), "complex128")
    W = random_array(np.random.default_rng(2002), (w, d, d, w), "complex128")
    Ab = random_array(np.random.default_rng(2003), (chi, d, chi), "complex128")
    tE, tA, tW, tAb = (Tensor(E, I("awb")).to_device(), Tensor(A, I("bsc")).to_device(),
                       Tensor(W, I("wstv")).to_device(), Tensor(Ab, I("ate")).to_device())
    got = _chain_cfg2(tE, tA, tW, tAb).to_host().data
    T = np.tensordot(E, A, axes=([2], [0]))                         # a w s c
    T2 = np.einsum("awsc,wstv->atvc", T, W, optimize=True)
    ref = np.tensordot(Ab, T2, axes=([0, 1], [0, 1]))               # e v c
    assert got.shape == (chi, w, chi)
    assert rel_frobenius(got, ref) <= 1e-12


def test_cfg1_full_size_properties():
    """configs[0] at dim 64 (4096^3 GEMM-equivalent), scrambled layout A[k,i,l,j]·B[n,l,m,k] → C[m,j,n,i].
    Size-independent checks: linearity in A, and a slab of C against the oracle."""
    n = 64
    rng = np.random.default_rng(1000)
    A1 = random_array(rng, (n, n, n, n), "complex128")
    A2 = random_array(rng, (n, n, n, n), "complex128")
    B = random_array(np.random.default_rng(1001), (n, n, n, n), "complex128")
    tB = Tensor(B, I("nlmk")).to_device()
    c1 = binary_einsum(Tensor(A1, I("kilj")).to_device(), tB, out=I("mjni")).to_host().data
    c2 = binary_einsum(Tensor(A2, I("kilj")).to_device(), tB, out=I("mjni")).to_host().data
    c12 = binary_einsum(Tensor(A1 + 2.0 * A2, I("kilj")).to_device(), tB, out=I("mjni")).to_host().data
    assert rel_frobenius(c12, c1 + 2.0 * c2) <= 1e-12
    # slab i = 5: C[m,j,n,5] = sum_{k,l} A[k,5,l,j] B[n,l,m,k]
    ref = np.einsum("klj,nlmk->mjn", A1[:, 5, :, :], B, optimize=True)
    assert rel_frobenius(c1[:, :, :, 5], ref) <= 1e-12


def test_cfg3_peps_batched_c64():
    """configs[2]: PEPS double-layer ComplexF32, D=8, χ=256, batch hyperindex β=8 (reduced β/χ for the full
    oracle compare here; the full-size run is a slab check)."""
    rng = np.random.default_rng(3000)
    chi, D, beta = 256, 8, 8
    A = random_array(rng, (chi, D, D, chi, beta), "complex64")          # l k b m z
    B = random_array(np.random.default_rng(3001), (chi, D, D, chi, beta), "complex64")  # m k q r z
    got = binary_einsum(Tensor(A, I("lkbmz")).to_device(), Tensor(B, I("mkqrz")).to_device(),
                        out=I("lbqrz")).to_host().data
    assert got.shape == (chi, D, D, chi, beta)
    z = 3
    ref = np.einsum("lkbm,mkqr->lbqr", A[..., z].astype(np.complex128), B[..., z].astype(np.complex128), optimize=True)
    assert rel_frobenius(got[..., z], ref.astype(np.complex64)) <= 1e-5


def test_stats_count_launches():
    h = _lib.Handle.get()
    h.reset_stats()
    a, ia, b, ib, ic = build_case(PARITY_CASES[3], "complex128", seed=1)
    contract(a, ia, b, ib, ic, path=mb.PATH_GETT_F64)
    s = h.stats()
    assert s["launches_gett_f64"] == 1 and s["launches_total"] >= 1


# ---- K3: tcgen05 / TMEM split-operand path ------------------------------------------------------------------------
TC_CASES = [
    ("tc_matmul_kmajor", dict(i=200, j=72, k=128), "ki", "kj", "ij"),
    ("tc_matmul_mmajor", dict(i=200, j=72, k=128), "ik", "jk", "ji"),
    ("tc_rank4", dict(a=16, b=12, c=8, d=24, e=10, f=16), "acbd", "dfce", "feab"),
    ("tc_batch_peps", dict(l=64, k=8, b=4, m=32, q=4, r=48, z=3), "lkbmz", "mkqrz", "lbqrz"),
    ("tc_multitile", dict(m=300, n=520, k=264, l=2), "kml", "nlk", "nml"),
    ("tc_multitile_T", dict(m=300, n=520, k=264, l=2), "kml", "nlk", "mln"),
    ("tc_rank8", dict(a=4, b=4, c=4, d=2, e=8, f=4, g=2, h=8, p=4, q=4, r=2, s=4), "aebfcgdh", "hpgqfres", "srqpdcba"),
    ("tc_skinny_n", dict(a=512, w=8, s=8, t=4, v=8), "aws", "wstv", "atv"),
]


def _tc_planned(case):
    name, ext, ia, ib, ic = case
    labels = {c: k for k, c in enumerate(dict.fromkeys(ia + ib))}
    h = _lib.Handle.get()
    return labels


@pytest.mark.parametrize("dt", ["complex64", "float32"])
@pytest.mark.parametrize("integer", [False, True], ids=["random", "integer_exact"])
@pytest.mark.parametrize("case", TC_CASES, ids=[c[0] for c in TC_CASES])
def test_tcgen05_split_parity(case, integer, dt):
    """ComplexF32 / Float32 through pack (K1 split writer) + tcgen05 GEMM, forced; <= 1e-5 rel. Frobenius against the
    oracle, and bit-exact on integer-valued inputs (hi parts exact, lo parts zero, fp32 accumulation exact)."""
    wide = np.complex128 if dt == "complex64" else np.float64
    a, ia, b, ib, ic = build_case(case, dt, seed=31, integer=integer)
    ref = binary_einsum_general(ic, a.astype(wide), ia, b.astype(wide), ib).astype(dt)
    h = _lib.Handle.get()
    h.reset_stats()
    got = contract(a, ia, b, ib, ic, device=True, path=mb.PATH_TCGEN05_TF32)
    s = h.stats()
    assert s["launches_tcgen05"] == 1 and s["launches_permute"] == 2, s   # the tcgen05 kernel really ran
    assert got.shape == ref.shape
    if integer:
        assert np.array_equal(got, ref), case[0]
    else:
        assert rel_frobenius(got, ref) <= 1e-5, (case[0], rel_frobenius(got, ref))


def test_tcgen05_accuracy_beats_plain_tf32():
    """Split-operand compensation (TF32 hi*hi + BF16 cross terms) + two-level accumulation must recover ~fp32 accuracy at long K
    (single-pass TF32 gives ~1e-3, a single TMEM accumulation chain drifts to 5.9e-5 at K = 4096)."""
    rng = np.random.default_rng(77)
    a = random_array(rng, (4096, 256), "complex64")
    b = random_array(rng, (4096, 192), "complex64")
    ref = (a.astype(np.complex128).T @ b.astype(np.complex128))
    got = contract(a, "ki", b, "kj", "ij", path=mb.PATH_TCGEN05_TF32)
    err = rel_frobenius(got.astype(np.complex128), ref)
    assert err <= 5e-6, err   # measured 1.3e-6 (3xTF32 scheme: 2.3e-6), flat in K thanks to the two-level accumulation
    # Float32: 128 x 256 tiles, one accumulator, K not a multiple of 16 (half-filled last smem line)
    ar = random_array(rng, (4104, 256), "float32")
    br = random_array(rng, (4104, 392), "float32")
    refr = ar.astype(np.float64).T @ br.astype(np.float64)
    h = _lib.Handle.get()
    h.reset_stats()
    gotr = contract(ar, "ki", br, "kj", "ij", path=mb.PATH_TCGEN05_TF32)
    assert h.stats()["launches_tcgen05"] == 1
    errr = rel_frobenius(gotr.astype(np.float64), refr)
    assert errr <= 5e-6, errr


@pytest.mark.parametrize("dt", ["complex64", "float32"])
def test_tcgen05_extreme_magnitudes(dt):
    """Operand values up to FLT_MAX (where rounding the hi / bf16 parts to nearest would overflow to infinity) against
    tiny ones, and a wide dynamic range inside one row: finite results within 1e-5 of the fp64 product."""
    rng = np.random.default_rng(23)
    K, M, N = 256, 128, 64
    a = random_array(rng, (K, M), dt)
    b = random_array(rng, (K, N), dt)
    scale_a = np.float32(2.0) ** rng.integers(100, 127, size=(K, 1)).astype(np.float32)    # up to ~1.7e38
    a = (a * scale_a).astype(dt)
    flat = a.reshape(-1).view(np.float32)
    flat[::97] = np.finfo(np.float32).max                                                 # FLT_MAX itself, both signs
    flat[5::193] = -np.finfo(np.float32).max
    b = (b * np.float32(2.0) ** -100).astype(dt)                                           # tiny partner: products stay finite
    wide = np.complex128 if dt == "complex64" else np.float64
    ref = a.astype(wide).T @ b.astype(wide)
    got = contract(np.asfortranarray(a), "ki", np.asfortranarray(b), "kj", "ij", path=mb.PATH_TCGEN05_TF32)
    assert np.isfinite(got).all()
    assert rel_frobenius(got.astype(wide), ref) <= 1e-5, rel_frobenius(got.astype(wide), ref)


@pytest.mark.parametrize("dt,m,n,k", [("complex64", 2048, 2048, 520), ("complex64", 1920, 2304, 264), ("float32", 2304, 2560, 392),
                                      ("float32", 4096, 4096, 128 * 3), ("complex64", 1300, 2000, 1032)])
def test_tcgen05_ragged_waves(dt, m, n, k):
    """Tile counts that leave a ragged last wave on the 148 persistent CTAs, ragged M / N edges and a partial last
    128-k chunk: integer-valued inputs must come out bit-exact."""
    rng = np.random.default_rng(5)
    a = integer_array(rng, (k, m), dt, lo=-2, hi=3)
    b = integer_array(rng, (k, n), dt, lo=-2, hi=3)
    wide = np.complex128 if dt == "complex64" else np.float64
    ref = (a.astype(wide).T @ b.astype(wide)).astype(dt)
    h = _lib.Handle.get()
    h.reset_stats()
    got = contract(a, "ki", b, "kj", "ij", path=mb.PATH_TCGEN05_TF32)
    assert h.stats()["launches_tcgen05"] == 1
    assert np.array_equal(got, ref)
    ar, br = random_array(rng, (k, m), dt), random_array(rng, (k, n), dt)
    got = contract(ar, "ki", br, "kj", "ij", path=mb.PATH_TCGEN05_TF32)
    assert rel_frobenius(got.astype(wide), ar.astype(wide).T @ br.astype(wide)) <= 1e-5


@pytest.mark.parametrize("dt", ["complex128", "float64", "complex64", "float32"])
@pytest.mark.parametrize("m,n,k,l", [(256, 256, 4096, 1), (64, 64, 8192, 1), (130, 70, 1000, 1), (512, 16, 2048, 1),
                                      (2048, 2, 2048, 1), (96, 96, 520, 3), (8, 300, 3000, 2)])
def test_gather_gemm_split_k(m, n, k, l, dt):
    """Too few tiles for 148 SMs -> the gather-GEMM runs every tile's k-range in slices (partial tiles to a workspace,
    ordered reduction pass). Integer-valued inputs must be bit-exact, ragged edges / batches / skinny tiles included."""
    rng = np.random.default_rng(8)
    a = integer_array(rng, (k, m, l), dt, lo=-2, hi=3)
    b = integer_array(rng, (n, l, k), dt, lo=-2, hi=3)
    ref = binary_einsum_general(list("nml"), a, list("kml"), b, list("nlk"))
    h = _lib.Handle.get()
    path = mb.PATH_GETT_F64 if dt in ("complex128", "float64") else mb.PATH_SIMT_F32
    got = contract(a, "kml", b, "nlk", "nml", device=True, path=path)
    assert np.array_equal(got, ref)
    ar, br = random_array(rng, (k, m, l), dt), random_array(rng, (n, l, k), dt)
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    refr = binary_einsum_general(list("nml"), ar.astype(wide), list("kml"), br.astype(wide), list("nlk"))
    got = contract(ar, "kml", br, "nlk", "nml", device=True, path=path)
    assert rel_frobenius(got.astype(wide), refr) <= (1e-12 if dt in ("complex128", "float64") else 1e-5)


F64_PAIR_CASES = [
    ("kmajor_matmul", dict(i=300, j=260, k=128), "ki", "kj", "ij"),                 # eligible: pairs of k are 16-byte elements
    ("kmajor_two_summed", dict(k=6, l=20, i=70, j=90), "klij", "klj", "ji"),         # k*l merged (120), C transposed (operands swap)
    ("kmajor_batch", dict(k=64, i=130, j=70, z=3), "kiz", "kjz", "ijz"),
    ("kmajor_split_k", dict(i=256, j=256, k=4096), "ki", "kj", "ij"),                # few tiles: split-K partial tiles in doubles
    ("k_odd_fallback", dict(i=200, j=136, k=129), "ki", "kj", "ij"),                 # odd K: plain Float64 core
    ("row_stride_odd_fallback", dict(k=64, i=131, j=70, z=3), "kiz", "jkz", "ijz"),  # B is not K-major
]


@pytest.mark.parametrize("integer", [False, True], ids=["random", "integer_exact"])
@pytest.mark.parametrize("case", F64_PAIR_CASES, ids=[c[0] for c in F64_PAIR_CASES])
def test_float64_k_pair_core(case, integer):
    """Float64 with both operands K-major takes the k-pair DMMA core (a 16-byte smem element = two consecutive k); odd K or a
    non-K-major operand fall back to the plain core. 1e-12 against the oracle, bit-exact on integer inputs."""
    a, ia, b, ib, ic = build_case(case, "float64", seed=43, integer=integer)
    ref = binary_einsum_general(ic, a, ia, b, ib)
    got = contract(a, ia, b, ib, ic, device=True, path=mb.PATH_GETT_F64)
    if integer:
        assert np.array_equal(got, ref), case[0]
    else:
        assert rel_frobenius(got, ref) <= 1e-12, (case[0], rel_frobenius(got, ref))


@pytest.mark.parametrize("m,n,k,l", [(2500, 1000, 160, 1), (1300, 700, 300, 3), (4096, 1200, 512, 1), (2500, 1000, 2048, 1)])
def test_gather_gemm_persistent_and_split_tail(m, n, k, l):
    """ComplexF64 shapes with several tiles per SM: short sums (K <= 512) take the persistent kernel (one software pipeline
    across tiles: the next tile's first k-block is gathered during the current tile's last one), a ragged last wave runs as
    a split tail launch (the K = 2048 case: 320 tiles = 296 + 24). Ragged M / N edges, batches; bit-exact on integer inputs."""
    dt = "complex128"
    rng = np.random.default_rng(10)
    a = integer_array(rng, (k, m, l), dt, lo=-2, hi=3)
    b = integer_array(rng, (n, l, k), dt, lo=-2, hi=3)
    ref = np.einsum("kml,nlk->nml", a, b)
    got = contract(a, "kml", b, "nlk", "nml", device=True, path=mb.PATH_GETT_F64)
    assert np.array_equal(got, ref)
    ar, br = random_array(rng, (k, m, l), dt), random_array(rng, (n, l, k), dt)
    got = contract(ar, "kml", br, "nlk", "mnl", device=True, path=mb.PATH_GETT_F64)
    assert rel_frobenius(got, np.einsum("kml,nlk->mnl", ar, br)) <= 1e-12


@pytest.mark.parametrize("dt", ["complex64", "float32"])
@pytest.mark.parametrize("m,n,k,l", [(256, 256, 4096, 1), (64, 64, 8192, 1), (130, 70, 1032, 1), (512, 512, 1024, 1), (96, 96, 520, 3)])
def test_tcgen05_split_k(m, n, k, l, dt):
    """The tcgen05 kernel's split-K schedule (few tiles: work units = (tile, k-slice), partial tiles through the
    workspace): bit-exact on integer-valued inputs, 1e-5 on random ones."""
    rng = np.random.default_rng(9)
    a = integer_array(rng, (k, m, l), dt, lo=-2, hi=3)
    b = integer_array(rng, (k, n, l), dt, lo=-2, hi=3)
    ref = binary_einsum_general(list("nml"), a, list("kml"), b, list("knl"))
    h = _lib.Handle.get()
    h.reset_stats()
    got = contract(a, "kml", b, "knl", "nml", device=True, path=mb.PATH_TCGEN05_TF32)
    assert h.stats()["launches_tcgen05"] == 1
    assert np.array_equal(got, ref)
    ar, br = random_array(rng, (k, m, l), dt), random_array(rng, (k, n, l), dt)
    wide = np.complex128 if dt == "complex64" else np.float64
    refr = binary_einsum_general(list("nml"), ar.astype(wide), list("kml"), br.astype(wide), list("knl"))
    got = contract(ar, "kml", br, "knl", "nml", device=True, path=mb.PATH_TCGEN05_TF32)
    assert rel_frobenius(got.astype(wide), refr) <= 1e-5


@pytest.mark.parametrize("dt,m,n,k,l", [("complex64", 4096, 4736, 264, 1), ("complex64", 1100, 2100, 520, 9), ("float32", 4608, 8448, 392, 1),
                                        ("float32", 3000, 1400, 136, 12)])
def test_tcgen05_cta_pair(dt, m, n, k, l):
    """Shapes with work for every SM take the CTA-pair kernel (cluster of 2, tcgen05 cta_group::2, 256-row tiles, each CTA
    holding half of the column operand): bit-exact on integer-valued inputs (ragged M / N edges, odd 128-row tile counts
    so that a pair's second CTA is entirely out of range, batches), 1e-5 on random ones."""
    rng = np.random.default_rng(12)
    a = integer_array(rng, (k, m, l), dt, lo=-2, hi=3)
    b = integer_array(rng, (k, n, l), dt, lo=-2, hi=3)
    wide = np.complex128 if dt == "complex64" else np.float64
    ref = np.einsum("kml,knl->nml", a.astype(wide), b.astype(wide)).astype(dt)
    h = _lib.Handle.get()
    h.reset_stats()
    got = contract(a, "kml", b, "knl", "nml", device=True, path=mb.PATH_TCGEN05_TF32)
    st = h.stats()
    assert st["launches_tcgen05"] == 1 and st["launches_tcgen05_pair"] == 1, st
    assert np.array_equal(got, ref)
    ar, br = random_array(rng, (k, m, l), dt), random_array(rng, (k, n, l), dt)
    refr = np.einsum("kml,knl->nml", ar.astype(wide), br.astype(wide))
    got = contract(ar, "kml", br, "knl", "mnl", device=True, path=mb.PATH_TCGEN05_TF32)
    assert rel_frobenius(got.astype(wide), refr.transpose(1, 0, 2)) <= 1e-5


def test_tcgen05_suite_with_cta_pairs_forced():
    """The whole tcgen05 parity battery once more with MB200_CTA_PAIR=2 (pair kernel for every shape with M > 128,
    however few tiles) — in a child process, the policy is read once per process."""
    import subprocess, sys
    env = dict(os.environ, MB200_CTA_PAIR="2")
    r = subprocess.run([sys.executable, "-m", "pytest", os.path.abspath(__file__), "-x", "-q", "-m", "gpu", "-k",
                        "tcgen05_split_parity or tcgen05_ragged or tcgen05_accuracy or qubit"],
                       env=env, capture_output=True, text=True, timeout=1200)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_tcgen05_auto_selected_for_large_c64():
    info = mb.plan_describe(_lib.C64, [0, 2, 4, 5, 6], _lib.C64, [0, 1, 2, 3, 6], [256, 8, 8, 256, 8],
                            _lib.C64, [3, 1, 4, 5, 6], [256, 8, 8, 256, 8])
    assert info.path == mb.PATH_TCGEN05_TF32
    info = mb.plan_describe(_lib.F32, [0, 2], _lib.F32, [0, 1], [4096, 4096], _lib.F32, [1, 2], [4096, 4096])
    assert info.path == mb.PATH_TCGEN05_TF32      # Float32 takes the tensor cores too (128 x 256 tiles)
    # summed extent not a multiple of 8: still the tensor cores (gather pack, K zero-padded) once the contraction is big enough
    info = mb.plan_describe(_lib.C64, [0, 2], _lib.C64, [1, 0], [100, 2048], _lib.C64, [1, 2], [100, 2048])
    assert info.path == mb.PATH_TCGEN05_TF32
    info = mb.plan_describe(_lib.C64, [0, 2], _lib.C64, [1, 0], [100, 512], _lib.C64, [1, 2], [100, 512])
    assert info.path == mb.PATH_SIMT_F32          # 2.6e7 MACs: below the tcgen05 size floor


TC_GATHER_CASES = [
    ("k100_kmajor", dict(i=300, j=200, k=100), "ki", "kj", "ij"),                    # K = 100 -> padded to 104
    ("k100_mmajor", dict(i=300, j=200, k=100), "ik", "jk", "ji"),                    # rows fastest in both operands
    ("bond_3_5_7", dict(a=3, b=5, c=7, i=130, j=70), "aibc", "cjba", "ji"),          # K = 105, three odd summed modes
    ("peps_d3", dict(l=27, k=9, b=3, m=27, q=9, r=27, z=3), "lkbmz", "mkqrz", "lbqrz"),   # D = 3 PEPS-like, batch z
    ("mixed_order", dict(a=6, b=10, i=96, j=40, l=4), "aibl", "bjla", "ijl"),       # K = 60 < 64 -> not eligible, FFMA fallback
]


@pytest.mark.parametrize("dt", ["complex64", "float32"])
@pytest.mark.parametrize("integer", [False, True], ids=["random", "integer_exact"])
@pytest.mark.parametrize("case", TC_GATHER_CASES, ids=[c[0] for c in TC_GATHER_CASES])
def test_tcgen05_gather_pack_parity(case, integer, dt):
    """Summed extents that do not tile groups of 8 k (K = 100, odd bond dimensions) take the table-driven gather pack (K
    zero-padded to a multiple of 8) and the same tcgen05 GEMM: <= 1e-5 against the oracle, bit-exact on integer inputs."""
    wide = np.complex128 if dt == "complex64" else np.float64
    a, ia, b, ib, ic = build_case(case, dt, seed=37, integer=integer)
    ref = binary_einsum_general(ic, a.astype(wide), ia, b.astype(wide), ib).astype(dt)
    h = _lib.Handle.get()
    h.reset_stats()
    got = contract(a, ia, b, ib, ic, device=True, path=mb.PATH_TCGEN05_TF32)
    s = h.stats()
    if case[0] == "mixed_order":
        assert s["launches_tcgen05"] == 0 and s["launches_simt_f32"] == 1, s
    else:
        assert s["launches_tcgen05"] == 1 and s["launches_permute"] == 2, s
    if integer:
        assert np.array_equal(got, ref), case[0]
    else:
        assert rel_frobenius(got, ref) <= 1e-5, (case[0], rel_frobenius(got, ref))


@pytest.mark.parametrize("dt", ["complex64", "float32"])
def test_tcgen05_strided_operands(dt):
    """Strided (non-dense) operands through the C ABI on the tcgen05 path: a slab of a larger array as the row operand
    (free-index shard) and every other row of a wider buffer as the column operand, strides passed explicitly."""
    rng = np.random.default_rng(41)
    a_full = random_array(rng, (192, 160, 6), dt)          # [i, k, s]; use i in 32:160, s in 2:5
    b_full = random_array(rng, (320, 96), dt)              # [2k, j]; use rows 0::2
    dA, dB = B200Array.from_host(a_full), B200Array.from_host(b_full)
    dC = B200Array((128, 96, 3), dt)
    h = _lib.Handle.get()
    h.set_path(mb.PATH_TCGEN05_TF32)
    h.reset_stats()
    esz = np.dtype(dt).itemsize
    en = _lib.dtype_enum(dt)
    try:
        _lib.check(mb.lib().mb200_binary_einsum(
            h.ptr, C.c_void_p(dC.ptr), en, 3, _lib.i32([0, 2, 3]), None,
            C.c_void_p(dA.ptr + (2 * 192 * 160 + 32) * esz), en, 3, _lib.i32([0, 1, 3]), _lib.i64([128, 160, 3]), _lib.i64([1, 192, 192 * 160]),
            C.c_void_p(dB.ptr), en, 2, _lib.i32([1, 2]), _lib.i64([160, 96]), _lib.i64([2, 320])))
    finally:
        h.set_path(mb.PATH_AUTO)
    assert h.stats()["launches_tcgen05"] == 1, h.stats()
    wide = np.complex128 if dt == "complex64" else np.float64
    ref = np.einsum("iks,kj->ijs", a_full[32:160, :, 2:5].astype(wide), b_full[0::2].astype(wide))
    assert rel_frobenius(dC.to_host().astype(wide), ref) <= 1e-5


# ---- fused contraction + reduce-scatter epilogue (single GPU: the "peers" are local buffers) -----------------
@pytest.mark.parametrize("dt,path,tol,dims", [("complex128", "gett", 1e-12, (256, 128, 1024)), ("complex64", "tcgen05", 1e-5, (256, 128, 1024)),
                                              ("complex64", "tcgen05_pair", 1e-5, (4096, 2048, 1024))])
def test_fused_scatter_epilogue_and_slot_reduce(dt, path, tol, dims):
    """mb200_binary_einsum_scatter + mb200_reduce_slots with 4 emulated ranks on one GPU: every rank contracts its
    K-slice, the epilogue routes each element to the owner's staging slot, the owners sum their slots. The union
    of the slabs must equal the unsliced contraction (what all_reduce(SUM) of the partials would give). The third case is
    large enough for the CTA-pair tcgen05 kernel (both CTAs of a pair scatter their half of the 256-row tile)."""
    nranks, (Mx, Nx, Kx) = 4, dims
    pair = path == "tcgen05_pair"
    path = "tcgen05" if pair else path
    rng = np.random.default_rng(17)
    a = random_array(rng, (Kx, Mx), dt)          # [k, i]
    b = random_array(rng, (Kx, Nx), dt)          # [k, j]
    ref = a.T @ b                                 # C[i, j]
    numel = Mx * Nx
    slab = numel // nranks
    assert slab & (slab - 1) == 0
    h = _lib.Handle.get()
    h.set_path(mb.PATH_TCGEN05_TF32 if path == "tcgen05" else mb.PATH_AUTO)
    L = mb.lib()
    staging = [B200Array((nranks * slab,), dt) for _ in range(nranks)]
    for s in staging:
        _lib.check(L.mb200_memset(h.ptr, C.c_void_p(s.ptr), 0xFF, s.nbytes))   # NaN-fill: every slot must be written
    arr = (C.c_void_p * nranks)(*[s.ptr for s in staging])
    kc = Kx // nranks
    h.reset_stats()
    for r in range(nranks):
        da = B200Array.from_host(a[r * kc:(r + 1) * kc, :])
        db = B200Array.from_host(b[r * kc:(r + 1) * kc, :])
        _lib.check(L.mb200_binary_einsum_scatter(
            h.ptr, _lib.dtype_enum(dt), 2, _lib.i32([1, 2]),
            C.c_void_p(da.ptr), _lib.dtype_enum(dt), 2, _lib.i32([0, 1]), _lib.i64(da.shape), None,
            C.c_void_p(db.ptr), _lib.dtype_enum(dt), 2, _lib.i32([0, 2]), _lib.i64(db.shape), None,
            arr, nranks, r, slab.bit_length() - 1))
    st = h.stats()
    assert (st["launches_tcgen05"] if path == "tcgen05" else st["launches_gett_f64"]) == nranks, st
    assert st["launches_tcgen05_pair"] == (nranks if pair else 0), st
    out = np.empty(numel, dtype=dt)
    for o in range(nranks):
        d = B200Array((slab,), dt)
        _lib.check(L.mb200_reduce_slots(h.ptr, C.c_void_p(d.ptr), C.c_void_p(staging[o].ptr), _lib.dtype_enum(dt), slab, nranks))
        out[o * slab:(o + 1) * slab] = d.to_host()
    got = out.reshape((Mx, Nx), order="F")
    assert rel_frobenius(got, ref.astype(dt)) <= tol
    # the flag barrier: every emulated rank signals epoch 7 into every flag array, then the waiting slot sum runs
    flags = [B200Array((nranks,), np.float32) for _ in range(nranks)]
    for f in flags:
        _lib.check(L.mb200_memset(h.ptr, C.c_void_p(f.ptr), 0, f.nbytes))
    farr = (C.c_void_p * nranks)(*[f.ptr for f in flags])
    for r in range(nranks):
        _lib.check(L.mb200_signal_peers(h.ptr, farr, nranks, r, 7))
    out2 = np.empty(numel, dtype=dt)
    for o in range(nranks):
        d = B200Array((slab,), dt)
        _lib.check(L.mb200_reduce_slots_wait(h.ptr, C.c_void_p(d.ptr), C.c_void_p(staging[o].ptr), _lib.dtype_enum(dt), slab, nranks,
                                             C.c_void_p(flags[o].ptr), 7))
        out2[o * slab:(o + 1) * slab] = d.to_host()
    assert np.array_equal(out2, out)
    assert np.array_equal(flags[0].to_host().view(np.int32), np.full(nranks, 7, np.int32))
    # a contraction that is not on a tensor-core path is refused (callers fall back to all_reduce)
    h.set_path(mb.PATH_AUTO)
    tiny_a, tiny_b = B200Array.from_host(a[:8, :16].copy(order="F")), B200Array.from_host(b[:8, :16].copy(order="F"))
    with pytest.raises(mb.ArgumentError):
        _lib.check(L.mb200_binary_einsum_scatter(
            h.ptr, _lib.dtype_enum(dt), 2, _lib.i32([1, 2]),
            C.c_void_p(tiny_a.ptr), _lib.dtype_enum(dt), 2, _lib.i32([0, 1]), _lib.i64((8, 16)), None,
            C.c_void_p(tiny_b.ptr), _lib.dtype_enum(dt), 2, _lib.i32([0, 2]), _lib.i64((8, 16)), None,
            arr, nranks, 0, 6))


# ---- tensor-network style operands: many dim-2 / dim-4 indices -------------------------------------------------
QUBIT_CASES = [
    # (name, rank-a labels, rank-b labels, out labels): every label has extent 2 unless listed in `ext`
    ("q_rank12_6shared", "abcdefghijkl", "ghijklmnopqr", "abcdefmnopqr", {}),
    ("q_rank14_scrambled", "kalbmcndgehf", "pgqhrkslmtnu", "abcdefpqrstu", {}),
    ("q_mixed_4_2", "abcdefgh", "efghijkl", "lkjidcba", dict(e=4, f=2, g=4, h=2, a=8, i=8)),
    ("q_batch", "abcdefgz", "efghijkz", "abcdhijkz", dict(z=3)),
]


@pytest.mark.parametrize("dt", ["complex128", "complex64"])
@pytest.mark.parametrize("case", QUBIT_CASES, ids=[c[0] for c in QUBIT_CASES])
def test_qubit_style_tensors(case, dt):
    """High-rank tensors with dim-2 indices (quantum-circuit contractions): ComplexF64 on the DMMA gather-GEMM,
    ComplexF32 through the tcgen05 path whose 8-k groups span several small summed modes."""
    name, ia, ib, ic, extd = case
    ext = {c: extd.get(c, 2) for c in set(ia + ib)}
    rng = np.random.default_rng(5)
    a = random_array(rng, tuple(ext[c] for c in ia), dt)
    b = random_array(rng, tuple(ext[c] for c in ib), dt)
    ref = binary_einsum_general(list(ic), a.astype(np.complex128), list(ia), b.astype(np.complex128), list(ib)).astype(dt)
    for path in ([mb.PATH_AUTO, mb.PATH_GETT_F64] if dt == "complex128" else [mb.PATH_AUTO, mb.PATH_TCGEN05_TF32, mb.PATH_SIMT_F32]):
        h = _lib.Handle.get()
        h.reset_stats()
        got = contract(a, ia, b, ib, ic, path=path)
        if path == mb.PATH_TCGEN05_TF32 and name != "q_batch":   # q_batch is below the tcgen05 size floor (K = 8)
            assert h.stats()["launches_tcgen05"] == 1, (name, h.stats())   # eligible: 2*2*2 / 4*2 tile the 8-k group
        assert got.shape == ref.shape
        assert rel_frobenius(got, ref) <= TOL[dt], (name, dt, path)


@pytest.mark.parametrize("dt", DTYPES)
def test_dot_like_split_k(dt):
    """Few outputs, long sums (norms, inner products, isometry checks — src/Tensor.jl:520-536): split-K direct kernel."""
    rng = np.random.default_rng(8)
    h = _lib.Handle.get()
    # <psi|phi>: full contraction of two rank-4 tensors, K = 32^4 = 1 048 576
    a = random_array(rng, (32, 32, 32, 32), dt)
    b = random_array(rng, (32, 32, 32, 32), dt)
    h.reset_stats()
    got = contract(a, "ijkl", b, "lkji", "")
    assert h.stats()["launches_direct"] == 1
    ref = np.einsum("ijkl,lkji->", a.astype(np.complex128), b.astype(np.complex128))
    assert got.shape == () and abs(complex(got) - ref) <= (1e-5 if dt in ("float32", "complex64") else 1e-12) * max(abs(ref), np.sqrt(a.size))
    # a few outputs: C[i] = sum_{jkl} A[i,j,k,l] B[l,k,j], M = 5
    a = random_array(rng, (5, 64, 64, 16), dt)
    b = random_array(rng, (16, 64, 64), dt)
    got = contract(a, "ijkl", b, "lkj", "i")
    ref = np.einsum("ijkl,lkj->i", a.astype(np.complex128), b.astype(np.complex128))
    assert rel_frobenius(got.astype(np.complex128), ref) <= TOL[dt]
    # integer data: exact
    ai, bi = integer_array(rng, (3, 4096, 4), dt), integer_array(rng, (4, 4096, 2), dt)
    got = contract(ai, "ikl", bi, "lkj", "ji")
    assert np.array_equal(got, binary_einsum_general(list("ji"), ai, list("ikl"), bi, list("lkj")))


@pytest.mark.parametrize("qs", [[7], [0], [3, 11], [0, 1], [18, 19], [2, 9, 15]], ids=lambda q: "q" + "_".join(map(str, q)))
@pytest.mark.parametrize("dt", ["complex128", "complex64"])
def test_gate_application(dt, qs):
    """1-, 2- and 3-qubit gates applied to a 20-qubit state (M = 2^17..2^19 rows, N = K = 2..8): the streaming apply
    kernel (tiny operators), the DMMA streaming kernel (8 x 8, ComplexF64); the qiskit CX check of
    test/python/qiskit.jl:33-41 in spirit. Checked against numpy.einsum."""
    nq = 20
    rng = np.random.default_rng(12)
    psi = random_array(rng, (2,) * nq, dt)
    k = len(qs)
    gate = random_array(rng, (2,) * (2 * k), dt)                # [o..., i...]
    labels = [f"q{i}" for i in range(nq)]
    outs = [f"o{i}" for i in range(k)]
    ia = outs + [labels[q] for q in qs]
    ic = list(labels)
    for o, q in zip(outs, qs):
        ic[q] = o
    tg = Tensor(gate, [Index(x) for x in ia]).to_device()
    tp = Tensor(psi, [Index(x) for x in labels]).to_device()
    h = _lib.Handle.get()
    h.reset_stats()
    got = binary_einsum(tg, tp, out=[Index(x) for x in ic]).to_host().data
    st = h.stats()
    if k == 3 and dt == "complex128":
        assert st["launches_gett_f64"] == 1, st                 # 8 x 8 operator: DMMA streaming kernel
    else:
        assert st["launches_direct"] == 1, st                   # apply kernel
    letters = "abcdefghijklmnopqrst"
    big = "XYZ"
    sub = letters
    for o, q in zip(big, qs):
        sub = sub.replace(letters[q], o)
    ref = np.einsum(f"{big[:k]}{''.join(letters[q] for q in qs)},{letters}->{sub}", gate.astype(np.complex128), psi.astype(np.complex128))
    assert rel_frobenius(got.astype(np.complex128), ref) <= TOL[dt]
    # and the same gate through the generic direct kernel must agree bit for bit in placement (integer data)
    gi, pi = integer_array(rng, gate.shape, dt), integer_array(rng, psi.shape, dt)
    a1 = contract(gi, ia, pi, labels, ic, path=mb.PATH_AUTO)
    _lib.Handle.get().set_path(mb.PATH_SIMT_F32 if dt == "complex64" else mb.PATH_GETT_F64)
    a2 = binary_einsum(BackendB200(), [Index(x) for x in ic], Tensor(gi, [Index(x) for x in ia]).to_device(),
                       Tensor(pi, [Index(x) for x in labels]).to_device()).to_host().data
    assert np.array_equal(a1, a2)
