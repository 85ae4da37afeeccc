"""`unary_einsum` / `hadamard`: the oracle is pinned on the reference's own known-answer tests
(CPU), then the CUDA path is compared with the oracle through the C ABI (GPU)."""
import numpy as np
import pytest

from cases import integer_array, random_array
from oracle import ArgumentError as OracleArgumentError
from oracle import hadamard_base, rel_frobenius, unary_einsum as oracle_unary, unary_einsum_general

DTYPES = ["float32", "float64", "complex64", "complex128"]
TOL = {"float32": 1e-5, "complex64": 1e-5, "float64": 1e-12, "complex128": 1e-12}

# test/integration/omeinsum.jl:6-100 (same battery, commented out, in test/unit/operations/unary_einsum.jl):
# (name, shape, inds, kwargs, expected inds, expected value (all-ones input))
UNARY_BATTERY = [
    ("axis_sum_dims", (2, 3, 4), "ijk", {"dims": "i"}, "jk", 2.0),        # :9-12
    ("axis_sum_out", (2, 3, 4), "ijk", {"out": "jk"}, "jk", 2.0),         # :14-15
    ("sum_all_dims", (2, 3, 4), "ijk", {"dims": "ijk"}, "", 24.0),        # :24-26
    ("sum_all_out", (2, 3, 4), "ijk", {"out": ""}, "", 24.0),             # :28-30
    ("diag_ii_i", (2, 2), "ii", {"out": "i"}, "i", 1.0),                  # :41-44
    ("diag_iji_ij", (2, 3, 2), "iji", {"out": "ij"}, "ij", 1.0),          # :52-55
    ("trace_default", (2, 2), "ii", {}, "", 2.0),                         # :66-68
    ("trace_dims", (2, 2), "ii", {"dims": "i"}, "", 2.0),                 # :70-72
    ("trace_out", (2, 2), "ii", {"out": ""}, "", 2.0),                    # :74-76
    ("ptrace_default", (2, 3, 2), "iji", {}, "j", 2.0),                   # :85-87
    ("ptrace_dims", (2, 3, 2), "iji", {"dims": "i"}, "j", 2.0),           # :89-91
    ("ptrace_out", (2, 3, 2), "iji", {"out": "j"}, "j", 2.0),             # :93-95
]

# random-data cases: (shape, inds_x, inds_y)
UNARY_CASES = [
    ((6, 5, 4), "ijk", "jk"), ((6, 5, 4), "ijk", "ik"), ((6, 5, 4), "ijk", "ij"), ((6, 5, 4), "ijk", "kji"),
    ((6, 5, 4), "ijk", ""), ((7, 7), "ii", "i"), ((7, 7), "ii", ""), ((5, 3, 5), "iji", "j"), ((5, 3, 5), "iji", "ji"),
    ((4, 3, 4, 3), "ijij", "ji"), ((4, 3, 4, 2), "ijik", "ki"), ((33, 17, 9), "abc", "ca"), ((33, 17, 9), "abc", "b"),
    ((64, 40), "ab", "b"), ((64, 40), "ab", "a"), ((300, 2, 3), "abc", "c"), ((1, 5, 1), "abc", "b"), ((3, 3, 3), "iii", "i"),
    ((3, 3, 3), "iii", ""), ((2,) * 10, "abcdefghij", "jfb"), ((16, 16, 16), "abc", "cab"),
]

# (shape_a, inds_a, shape_b, inds_b)
HADAMARD_CASES = [
    ((2, 3, 4), "ijk", (), ""), ((2, 3, 4), "ijk", (2,), "i"), ((2, 3, 4), "ijk", (3,), "j"), ((2, 3, 4), "ijk", (4,), "k"),
    ((2, 3, 4), "ijk", (2, 3), "ij"), ((2, 3, 4), "ijk", (3, 2), "ji"), ((2, 3, 4), "ijk", (4, 2), "ki"),
    ((2, 3, 4), "ijk", (2, 3, 4), "ijk"), ((2, 3, 4), "ijk", (4, 3, 2), "kji"), ((8, 6, 4), "ijk", (8,), "i"),
    ((8, 6, 4), "ijk", (6, 8), "ji"), ((8, 6, 4), "ijk", (4,), "k"), ((16, 5), "ab", (16, 5), "ab"), ((7, 5), "ab", (7,), "a"),
    ((64, 33, 3), "abc", (3, 64), "ca"), ((1, 4, 1), "abc", (4,), "b"), ((12,), "a", (12,), "a"),
]


def _ix(s):
    return list(s)


# ---------------------------------------------------------------------------------------------------- CPU
@pytest.mark.parametrize("case", UNARY_BATTERY, ids=[c[0] for c in UNARY_BATTERY])
def test_oracle_unary_reference_known_answers(case):
    name, shape, inds, kw, exp_inds, exp_val = case
    for dt in (np.float64, np.complex128):
        x = np.ones(shape, dt)
        y, inds_y = oracle_unary(x, _ix(inds), **{k: _ix(v) for k, v in kw.items()})
        assert inds_y == _ix(exp_inds)
        assert y.shape == tuple(shape[inds.index(i)] for i in exp_inds)
        assert np.array_equal(y, np.full(y.shape, exp_val, dt))


def test_oracle_unary_matches_numpy_einsum_and_rejects():
    rng = np.random.default_rng(0)
    for shape, ix, iy in UNARY_CASES:
        x = random_array(rng, shape, "complex128")
        m = {c: k for k, c in enumerate(dict.fromkeys(ix))}
        ref = np.einsum(x, [m[c] for c in ix], [m[c] for c in iy])
        assert rel_frobenius(unary_einsum_general(_ix(iy), x, _ix(ix)), ref) < 1e-14
    with pytest.raises(OracleArgumentError):      # ext/MuscleOMEinsumExt.jl:32
        unary_einsum_general(_ix("iz"), np.ones((2, 3)), _ix("ij"))


def test_oracle_hadamard_reference_known_answers():
    """test/unit/operations/hadamard.jl:4-139."""
    a = np.ones((2, 3, 4))
    c, inds = hadamard_base(a, _ix("ijk"), np.array(2.0), [])                       # :4-20
    assert inds == _ix("ijk") and np.array_equal(c, 2.0 * a)
    for lab, n in (("i", 2), ("j", 3), ("k", 4)):                                   # :22-81
        b = np.arange(1.0, n + 1)
        c, inds = hadamard_base(a, _ix("ijk"), b, [lab])
        assert inds == _ix("ijk")
        for d in range(n):
            assert np.all(np.take(c, d, axis="ijk".index(lab)) == d + 1)
    b = np.array([[1.0, 2, 3], [4, 5, 6]])                                           # :83-108
    c, inds = hadamard_base(a, _ix("ijk"), b, _ix("ij"))
    for i in range(2):
        for j in range(3):
            assert np.all(c[i, j, :] == b[i, j])
    c, _ = hadamard_base(a, _ix("ijk"), 2 * np.ones((2, 3, 4)), _ix("ijk"))        # :110-126
    assert np.all(c == 2.0)
    # lower-rank operand first: swapped (hadamard.jl:8)
    c, inds = hadamard_base(np.arange(1.0, 4), ["j"], a, _ix("ijk"))
    assert inds == _ix("ijk") and np.all(c[:, 2, :] == 3.0)
    with pytest.raises(OracleArgumentError):                                         # :10
        hadamard_base(a, _ix("ijk"), np.ones(5), ["z"])


def test_frontend_and_dispatch_without_gpu():
    import muscle_b200 as mb
    from muscle_b200 import ArgumentError, Index, Tensor
    I = lambda s: [Index(c) for c in s]
    assert mb.unary_frontend_inds_y(I("iji")) == I("j")
    assert mb.unary_frontend_inds_y(I("ijk"), dims=I("i")) == I("jk")
    assert mb.unary_frontend_inds_y(I("ijk"), out=I("kj")) == I("kj")
    assert mb.unary_frontend_inds_y(I("ii"), out=I("i")) == I("i")
    # host arrays keep the reference's backends, which are not part of this package: "not implemented or not loaded"
    x = Tensor(np.ones((2, 2)), I("ii"))
    with pytest.raises(ArgumentError, match="not implemented or not loaded"):
        mb.unary_einsum(x)
    with pytest.raises(ArgumentError, match="not implemented or not loaded"):
        mb.hadamard(Tensor(np.ones((2, 3)), I("ij")), Tensor(np.ones(2), I("i")))
    with pytest.raises(ArgumentError):    # hadamard.jl:10
        mb.hadamard(Tensor(np.ones((2, 3)), I("ij")), Tensor(np.ones(2), I("z")))


# ---------------------------------------------------------------------------------------------------- GPU
def _dev(x, inds):
    from muscle_b200 import Index, Tensor
    return Tensor(x, [Index(c) for c in inds]).to_device()


@pytest.mark.gpu
@pytest.mark.parametrize("case", UNARY_BATTERY, ids=[c[0] for c in UNARY_BATTERY])
def test_unary_reference_battery_on_b200(case):
    import muscle_b200 as mb
    from muscle_b200 import BackendB200, Index, Tensor
    name, shape, inds, kw, exp_inds, exp_val = case
    I = lambda s: [Index(c) for c in s]
    for dt in DTYPES:
        kwargs = {k: I(v) for k, v in kw.items()}
        x = Tensor(np.ones(shape, dt), I(inds))
        y = mb.unary_einsum(x.to_device(), **kwargs)                       # Domain(B200Array) selects BackendB200
        assert y.inds == I(exp_inds) and y.on_device
        got = y.to_host().data
        assert got.dtype == np.dtype(dt) and np.array_equal(got, np.full(got.shape, exp_val, dt))
        yh = mb.with_backend(lambda: mb.unary_einsum(x, **kwargs), BackendB200())   # host arrays through the same .so
        assert not yh.on_device and np.array_equal(yh.data, got)
        out = Tensor(np.zeros(got.shape, dt), I(exp_inds)).to_device()      # unary_einsum!(B, A): omeinsum.jl:17-19
        assert mb.unary_einsum_(out, x.to_device()) is out
        assert np.array_equal(out.to_host().data, got)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,ix,iy", UNARY_CASES)
def test_unary_parity(shape, ix, iy, dt):
    import muscle_b200 as mb
    from muscle_b200 import Index
    rng = np.random.default_rng(3)
    x = random_array(rng, shape, dt)
    got = mb.unary_einsum(_dev(x, ix), out=[Index(c) for c in iy]).to_host().data
    ref = unary_einsum_general(_ix(iy), x.astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64), _ix(ix))
    assert got.shape == ref.shape
    assert rel_frobenius(got.astype(ref.dtype), ref) <= TOL[dt]
    xi = integer_array(rng, shape, dt)                                       # exact in any order: bookkeeping ==
    goti = mb.unary_einsum(_dev(xi, ix), out=[Index(c) for c in iy]).to_host().data
    assert np.array_equal(goti, unary_einsum_general(_ix(iy), xi, _ix(ix)))


@pytest.mark.gpu
def test_unary_large_forms():
    """The three kernel forms at sizes that fill the GPU: thread-per-output (kept unit-stride mode), warp-per-output
    (summed unit-stride mode), split (few outputs)."""
    import muscle_b200 as mb
    from muscle_b200 import Index
    rng = np.random.default_rng(5)
    for dt in ("complex128", "float32"):
        x = integer_array(rng, (512, 96, 40), dt, lo=-2, hi=3)
        for iy in ("ac", "bc", "c", "", "ca", "b"):
            got = mb.unary_einsum(_dev(x, "abc"), out=[Index(c) for c in iy]).to_host().data
            assert np.array_equal(got, unary_einsum_general(_ix(iy), x, _ix("abc"))), (dt, iy)
        d = integer_array(rng, (300, 7, 300), dt)
        for iy in ("ab", "b", "", "ba"):
            got = mb.unary_einsum(_dev(d, "aba"), out=[Index(c) for c in iy]).to_host().data
            assert np.array_equal(got, unary_einsum_general(_ix(iy), d, _ix("aba"))), (dt, iy)
    s = mb.Handle.get(0).stats()
    assert s["launches_unary"] > 0


@pytest.mark.gpu
def test_unary_rejects():
    import muscle_b200 as mb
    from muscle_b200 import ArgumentError, Index
    x = _dev(np.ones((2, 3)), "ij")
    with pytest.raises(ArgumentError, match="subset"):
        mb.unary_einsum(x, out=[Index("i"), Index("z")])
    with pytest.raises(ArgumentError):
        mb.unary_einsum(x, out=[Index("i"), Index("i")])


@pytest.mark.gpu
def test_hadamard_reference_battery_on_b200():
    """test/unit/operations/hadamard.jl:4-139 on BackendB200, out-of-place and `hadamard!(a, a, b)`."""
    import muscle_b200 as mb
    from muscle_b200 import BackendB200, Index, Tensor
    I = lambda s: [Index(c) for c in s]
    for dt in DTYPES:
        ones = np.ones((2, 3, 4), dt)
        cases = [(np.array(2.0, dt), ""), (np.arange(1, 3).astype(dt), "i"), (np.arange(1, 4).astype(dt), "j"),
                 (np.arange(1, 5).astype(dt), "k"), (np.array([[1, 2, 3], [4, 5, 6]]).astype(dt), "ij"),
                 (2 * np.ones((2, 3, 4), dt), "ijk")]
        for b, ib in cases:
            ref, _ = hadamard_base(ones, _ix("ijk"), b, _ix(ib))
            a = Tensor(ones.copy(), I("ijk")).to_device()
            bt = Tensor(b, I(ib)).to_device()
            c = mb.hadamard(a, bt)
            assert c.inds == I("ijk") and c.shape == (2, 3, 4) and np.array_equal(c.to_host().data, ref)
            assert np.array_equal(mb.hadamard(bt, a).to_host().data, ref)          # swapped operands (hadamard.jl:8)
            r = mb.hadamard_(a, a, bt)                                              # c === a
            assert r is a and np.array_equal(a.to_host().data, ref)
            ch = mb.with_backend(lambda: mb.hadamard(Tensor(ones, I("ijk")), Tensor(b, I(ib))), BackendB200())
            assert not ch.on_device and np.array_equal(ch.data, ref)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("sa,ia,sb,ib", HADAMARD_CASES)
def test_hadamard_parity(sa, ia, sb, ib, dt):
    import muscle_b200 as mb
    rng = np.random.default_rng(7)
    a, b = random_array(rng, sa, dt), random_array(rng, sb, dt)
    got = mb.hadamard(_dev(a, ia), _dev(b, ib)).to_host().data
    ref, _ = hadamard_base(a, _ix(ia), b, _ix(ib))
    assert got.shape == ref.shape and got.dtype == ref.dtype
    assert rel_frobenius(got, ref) <= TOL[dt]
    ai, bi = integer_array(rng, sa, dt), integer_array(rng, sb, dt)
    goti = mb.hadamard(_dev(ai, ia), _dev(bi, ib)).to_host().data
    assert np.array_equal(goti, hadamard_base(ai, _ix(ia), bi, _ix(ib))[0])


@pytest.mark.gpu
def test_hadamard_mixed_eltypes_and_large():
    import muscle_b200 as mb
    rng = np.random.default_rng(9)
    for da, db in (("float64", "complex128"), ("complex64", "float32"), ("float32", "float64"), ("complex128", "float64")):
        a, b = integer_array(rng, (6, 5, 4), da), integer_array(rng, (4, 6), db)
        got = mb.hadamard(_dev(a, "ijk"), _dev(b, "ki")).to_host().data
        ref, _ = hadamard_base(a, _ix("ijk"), b, _ix("ki"))
        assert got.dtype == ref.dtype and np.array_equal(got, ref)
    for dt in ("complex64", "float32", "complex128"):
        a = integer_array(rng, (1024, 48, 20), dt)
        for sb, ib in (((1024,), "a"), ((48,), "b"), ((20, 1024), "ca"), ((1024, 48, 20), "abc"), ((48, 20), "bc")):
            b = integer_array(rng, sb, dt)
            got = mb.hadamard(_dev(a, "abc"), _dev(b, ib)).to_host().data
            assert np.array_equal(got, hadamard_base(a, _ix("abc"), b, _ix(ib))[0]), (dt, ib)


@pytest.mark.gpu
def test_unary_strided_operands_through_the_c_abi():
    """The ABI takes element strides for x and y (NULL = dense): reduce a strided sub-block of a larger array into a
    strided destination, compare with the oracle on the same views."""
    import ctypes as C
    import muscle_b200 as mb
    from muscle_b200 import B200Array, _lib
    rng = np.random.default_rng(13)
    for dt in ("complex128", "float32"):
        big = integer_array(rng, (10, 9, 8), dt)
        out = np.zeros((7, 12), dt, order="F")
        X = B200Array.from_host(big)
        Y = B200Array.from_host(out)
        # x = big[2:8, 1:8:2, 3:7]  (6, 4, 4) with element strides (1, 20, 90), offset 2 + 10 + 270; y = out[1:7, ::3][:, :4]
        xs, ys = (1, 20, 90), (1, 21)
        xoff, yoff = (2 + 1 * 10 + 3 * 90) * big.itemsize, 1 * out.itemsize
        h = _lib.Handle.get()
        _lib.check(mb.lib().mb200_unary_einsum(
            h.ptr, C.c_void_p(Y.ptr + yoff), _lib.dtype_enum(dt), 2, _lib.i32([0, 2]), _lib.i64(ys),
            C.c_void_p(X.ptr + xoff), _lib.dtype_enum(dt), 3, _lib.i32([0, 1, 2]), _lib.i64((6, 4, 4)), _lib.i64(xs)))
        got = Y.to_host()
        ref = out.copy()
        ref[1:7, 0:12:3] = unary_einsum_general(list("ac"), big[2:8, 1:8:2, 3:7], list("abc"))
        assert np.array_equal(got, ref)
