"""CPU: pins the oracle (oracle/) against the reference's own known-answer tests, against the
independent plain-C loop nest, and against the committed golden vectors."""
import os

import numpy as np
import pytest

from cases import PARITY_CASES, REFERENCE_BATTERY, build_case
from oracle import (ArgumentError, DimensionMismatch, binary_einsum, binary_einsum_base,
                    binary_einsum_base_inplace, binary_einsum_general, frontend_inds_c, rel_frobenius)
from oracle.build_oracle import einsum_loops

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "binary_einsum_golden.npz")


@pytest.mark.parametrize("case", REFERENCE_BATTERY, ids=[c[0] for c in REFERENCE_BATTERY])
def test_reference_known_answers(case):
    """test/unit/operations/binary_einsum.jl:4-154 (+ omeinsum.jl / cuda.jl for hyperindex cases)."""
    name, sa, ia, sb, ib, kw, exp_inds, exp_shape, exp_val, hyper = case
    for dt in (np.float64, np.complex128):
        a, b = np.ones(sa, dt), np.ones(sb, dt)
        kwargs = {k: list(v) for k, v in kw.items()}
        if hyper:
            # Muscle's default host backend rejects hyperindices (binary_einsum.jl:82-83; tests :122-133,153)
            with pytest.raises(ArgumentError):
                binary_einsum(a, list(ia), b, list(ib), general=False, **kwargs)
        c, inds_c = binary_einsum(a, list(ia), b, list(ib), general=True, **kwargs)
        assert inds_c == list(exp_inds)
        assert c.shape == exp_shape
        assert np.array_equal(c, np.full(exp_shape, exp_val, dt))


def test_reference_scale():
    """scale — test/unit/operations/binary_einsum.jl:96-119; mixed eltypes omeinsum.jl:160-186."""
    for dt in (np.float64, np.complex128):
        a = np.ones((2, 3), dt)
        alpha = np.array(2.0)
        for (x, ix, y, iy) in ((a, "ij", alpha, ""), (alpha, "", a, "ij")):
            c, inds = binary_einsum(x, list(ix), y, list(iy))
            assert inds == list("ij") and np.array_equal(c, 2.0 * a) and c.dtype == dt
            c, inds = binary_einsum(x, list(ix), y, list(iy), out=list("ji"))
            assert inds == list("ji") and np.array_equal(c, 2.0 * a.T)


def test_reference_manual_matches_reshape_matmul():
    """"manual" testset — test/unit/operations/binary_einsum.jl:135-149."""
    rng = np.random.default_rng(0)
    for dt in (np.float64, np.complex128):
        A = rng.normal(size=(2, 3, 4)).astype(dt)
        B = rng.normal(size=(4, 5, 3)).astype(dt)
        c, inds = binary_einsum(A, list("ijk"), B, list("klj"), dims=list("jk"))
        a_mat = np.reshape(A, (2, 12), order="F")
        b_mat = np.reshape(np.transpose(B, (2, 0, 1)), (12, 5), order="F")
        assert inds == list("il")
        assert rel_frobenius(c, a_mat @ b_mat) < 1e-14
        # hyperindex variant vs the explicit 4-nested loop (omeinsum.jl:225-236)
        c2 = binary_einsum_general(list("ikl"), A, list("ijk"), B, list("klj"))
        ref = np.zeros((2, 4, 5), dt)
        for i in range(2):
            for j in range(3):
                for k in range(4):
                    for l in range(5):
                        ref[i, k, l] += A[i, j, k] * B[k, l, j]
        assert rel_frobenius(c2, ref) < 1e-14


def test_frontend_semantics():
    """binary_einsum.jl:33-41."""
    assert frontend_inds_c(list("ij"), list("jk")) == list("ik")
    assert frontend_inds_c(list("ijb"), list("jkb"), dims=list("j")) == list("ibk")  # a's order, then b's
    assert frontend_inds_c(list("ij"), list("jk"), dims=list("jz")) == list("ik")    # dims ∩ a ∩ b
    assert frontend_inds_c(list("ij"), list("jk"), out=list("ki")) == list("ki")
    assert frontend_inds_c(list("ij"), list("ji"), dims=list("ji")) == []


def test_inplace_base_requires_left_right_order():
    """binary_einsum!(::BackendBase) — binary_einsum.jl:98-121 (`inds(c) == [left; right]`, :108)."""
    a, b = np.ones((2, 3)), np.ones((3, 4))
    c = np.zeros((2, 4))
    binary_einsum_base_inplace(c, list("ik"), a, list("ij"), b, list("jk"))
    assert np.array_equal(c, 3 * np.ones((2, 4)))
    with pytest.raises(ArgumentError):
        binary_einsum_base_inplace(np.zeros((4, 2)), list("ki"), a, list("ij"), b, list("jk"))


def test_tensor_ctor_checks():
    from oracle.muscle_oracle import check_tensor
    with pytest.raises(ArgumentError):
        check_tensor(np.ones((2, 3)), list("i"))
    with pytest.raises(DimensionMismatch):
        check_tensor(np.ones((2, 3)), list("ii"))


@pytest.mark.parametrize("case", PARITY_CASES, ids=[c[0] for c in PARITY_CASES])
@pytest.mark.parametrize("dt", ["float64", "complex128", "complex64"])
def test_numpy_oracle_vs_c_loop_nest(case, dt):
    """The BLAS/TTGT restatement and the explicit loop nest are independent; they must agree."""
    if int(np.prod(list(case[1].values()), dtype=np.int64)) > 3_000_000:
        pytest.skip("loop nest too slow for this case")
    a, ia, b, ib, ic = build_case(case, dt, seed=7)
    c = binary_einsum_general(ic, a, ia, b, ib)
    try:
        c_base = binary_einsum_base(ic, a, ia, b, ib)
        assert rel_frobenius(c_base, c) < (1e-5 if dt == "complex64" else 1e-13)
    except ArgumentError:
        pass  # hyperindex case: BackendBase rejects it, the general path is the semantics
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    ref = einsum_loops(ic, a.astype(wide), ia, b.astype(wide), ib)
    assert c.shape == ref.shape
    assert rel_frobenius(c, ref) < (1e-5 if dt == "complex64" else 1e-13)


def test_golden_vectors():
    g = np.load(GOLDEN)
    keys = sorted({k.rsplit("__", 1)[0] for k in g.files})
    assert len(keys) == 36
    by_name = {c[0]: c for c in PARITY_CASES}
    for key in keys:
        name, dt = key.split("__")
        _, _, ia, ib, ic = by_name[name]
        a, b, c = g[key + "__a"], g[key + "__b"], g[key + "__c"]
        got = binary_einsum_general(list(ic), a, list(ia), b, list(ib))
        tol = 1e-5 if dt in ("float32", "complex64") else 1e-12
        assert got.shape == c.shape and got.dtype == c.dtype
        assert rel_frobenius(got, c) <= tol, key


def test_integer_inputs_are_exact():
    """Integer-valued data makes every product exact: placement / bookkeeping can be checked with ==."""
    from cases import integer_array
    rng = np.random.default_rng(3)
    a = integer_array(rng, (5, 4, 3), np.complex128)
    b = integer_array(rng, (3, 6, 4), np.complex128)
    c1 = binary_einsum_base(list("li"), a, list("ijk"), b, list("klj"))
    c2 = einsum_loops(list("li"), a, list("ijk"), b, list("klj"))
    assert np.array_equal(c1, c2)


def test_dangling_label_semantics_of_the_two_oracles():
    """A label carried by one operand only and absent from the output: BackendBase throws (binary_einsum.jl:82-83) while the general
    einsum - OMEinsum / cuTENSOR semantics (ext/MuscleOMEinsumExt.jl:40-59, ext/MuscleCUDAExt.jl:30-38), the ones BackendB200
    replaces - sums it. The GPU tests (tests/test_at_size.py::test_dangling_*) check the CUDA path against the second."""
    import oracle
    rng = np.random.default_rng(4)
    a = rng.standard_normal((3, 5, 4))      # i x j
    b = rng.standard_normal((4, 6))         # j k
    with pytest.raises(oracle.ArgumentError):
        oracle.binary_einsum_base(list("ik"), a, list("ixj"), b, list("jk"))
    got = oracle.binary_einsum_general(list("ik"), a, list("ixj"), b, list("jk"))
    assert np.allclose(got, a.sum(axis=1) @ b)


def test_dangling_golden_vectors_pin_the_general_oracle():
    """tests/golden/dangling_golden.npz (plain-C loop nest, tests/golden/make_golden_dangling.py): the NumPy general oracle sums dangling
    labels exactly as the explicit loop does."""
    import importlib.util
    import oracle
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location("mgd", os.path.join(here, "golden", "make_golden_dangling.py"))
    mgd = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mgd)
    z = np.load(os.path.join(here, "golden", "dangling_golden.npz"))
    for name, ext, ia, ib, ic in mgd.CASES:
        for dt in mgd.DTYPES:
            a, b, c = z[f"{name}__{dt}__a"], z[f"{name}__{dt}__b"], z[f"{name}__{dt}__c"]
            wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
            got = oracle.binary_einsum_general(list(ic), a.astype(wide), list(ia), b.astype(wide), list(ib))
            tol = 1e-12 if dt in ("float64", "complex128") else 1e-6
            assert got.shape == c.shape and oracle.rel_frobenius(got, c.astype(wide)) <= tol, (name, dt)
