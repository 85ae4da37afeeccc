"""`tensor_svd_thin` / `simple_update`. Singular vectors are unique only up to phases (and any
rotation inside a degenerate subspace), so parity is checked through the invariants the reference's own tests use
(test/unit/operations/tensor_svd_thin.jl:24-38, simple_update.jl:33-70): index/shape bookkeeping, singular values
against LAPACK (the oracle), reconstruction A = U·s·Vt, isometry of U and Vt, descending order."""
import numpy as np
import pytest

from cases import random_array
from oracle import rel_frobenius, simple_update_theta, tensor_svd_thin_base

DTYPES = ["float32", "float64", "complex64", "complex128"]
TOL = {"float32": 2e-5, "complex64": 2e-5, "float64": 1e-12, "complex128": 1e-12}

SVD_SHAPES = [((2, 4, 6, 8), "ijkl", "ij"), ((2, 4, 6, 8), "ijkl", "ikl"), ((2, 4, 6, 8), "ijkl", "l"), ((2, 4, 6, 8), "ijkl", "kj"),
              ((64, 64), "ab", "a"), ((200, 37), "ab", "a"), ((37, 200), "ab", "a"), ((1, 5), "ab", "a"), ((5, 1), "ab", "a"),
              ((33, 33), "ab", "b"), ((16, 2, 16, 2), "lpqr", "lp")]


def _gamma():
    # test/unit/operations/simple_update.jl:8-31: MPS factorisation of |00>+|01>+|10>+|11>, identity and CX gates
    ga = np.eye(2)
    gb = np.eye(2)
    ident = np.reshape(np.eye(4), (2, 2, 2, 2), order="F")
    cx = np.reshape(np.array([[1.0, 0, 0, 0], [0, 1, 0, 0], [0, 0, 0, 1], [0, 0, 1, 0]]), (2, 2, 2, 2), order="F")
    return ga, gb, ident, cx


# ------------------------------------------------------------------------------------------------ CPU
def test_factorinds_rules():
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    assert mb.factorinds(I("ijkl"), I("ij"), []) == (I("ij"), I("kl"))
    assert mb.factorinds(I("ijkl"), [], I("l")) == (I("ijk"), I("l"))
    assert mb.factorinds(I("ijkl"), I("ki"), I("jl")) == (I("ki"), I("jl"))
    for lu, lv in ((I("z"), []), ([], I("z")), (I("ijkl"), []), ([], I("ijkl")), (I("ij"), I("jk"))):
        with pytest.raises(mb.ArgumentError):                   # tensor_svd_thin.jl:10-18
            mb.factorinds(I("ijkl"), lu, lv)
    A = mb.Tensor(np.ones((2, 4, 6, 8)), I("ijkl"))
    with pytest.raises(mb.ArgumentError):                       # :9 (no inds given), host tensor -> reference backend
        mb.tensor_svd_thin(A)


def test_oracle_svd_and_simple_update_reference_known_answers():
    rng = np.random.default_rng(0)
    a = random_array(rng, (2, 4, 6, 8), "complex128")
    U, s, Vt = tensor_svd_thin_base(a, list("ijkl"), list("ij"), list("kl"))
    assert U.shape == (2, 4, 8) and s.shape == (8,) and Vt.shape == (6, 8, 8)          # tensor_svd_thin.jl:28-30
    assert rel_frobenius(np.einsum("ijx,x,klx->ijkl", U, s, Vt), a) < 1e-13             # :32
    assert np.allclose(np.einsum("ijx,ijy->xy", U.conj(), U), np.eye(8))                 # :33 isisometry
    ga, gb, ident, cx = _gamma()
    for g, sv in ((ident, [1.0, 1.0]), (cx, [np.sqrt(2.0), 0.0])):                      # simple_update.jl:47, :68
        th, ith = simple_update_theta(ga, ["pa", "bond"], gb, ["pb", "bond"], g, ["pa", "pb", "ga", "gb"],
                                      "pa", "pb", "bond", "ga", "gb")
        assert ith == ["pa", "pb"]
        _, s, _ = tensor_svd_thin_base(th, ith, ["pa"], ["pb"])
        assert np.allclose(s, sv, atol=1e-15)
        assert np.isclose(np.sum(np.abs(th) ** 2), 2.0)                                  # :49-50, :70-71


# ------------------------------------------------------------------------------------------------ GPU
def _check_svd(mb, a, inds, inds_u, dt):
    I = lambda s: [mb.Index(c) for c in s]
    A = mb.Tensor(a, I(inds)).to_device()
    U, S, Vt = mb.tensor_svd_thin(A, inds_u=I(inds_u), ind_s=mb.Index("x"))
    inds_v = [c for c in inds if c not in inds_u]
    assert U.inds == I(inds_u) + [mb.Index("x")] and S.inds == [mb.Index("x")] and Vt.inds == I(inds_v) + [mb.Index("x")]
    u, s, vt = U.to_host().data, S.to_host().data, Vt.to_host().data
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    Uo, so, Vto = tensor_svd_thin_base(a.astype(wide), list(inds), list(inds_u), inds_v)
    assert u.shape == Uo.shape and s.shape == so.shape and vt.shape == Vto.shape
    assert s.dtype == (np.float32 if dt in ("float32", "complex64") else np.float64)
    assert np.all(np.diff(s) <= 0) and np.all(s >= 0)                       # descending, non-negative
    assert np.linalg.norm(s - so) <= TOL[dt] * max(np.linalg.norm(so), 1e-300)
    k = s.shape[0]
    um, vm = u.reshape(-1, k, order="F").astype(wide), vt.reshape(-1, k, order="F").astype(wide)
    ref = np.transpose(a, [inds.index(c) for c in list(inds_u) + inds_v]).reshape(um.shape[0], vm.shape[0], order="F")
    rec = (um * s.astype(wide)) @ vm.T
    assert rel_frobenius(rec, ref.astype(wide)) <= TOL[dt]                  # A = U s Vt
    keep = s > 1e-3 * s[0] if s[0] > 0 else np.zeros(k, bool)               # isometry on the numerically non-null part
    gu = um[:, keep].conj().T @ um[:, keep]
    gv = vm[:, keep].conj().T @ vm[:, keep]
    assert np.linalg.norm(gu - np.eye(keep.sum())) <= 50 * TOL[dt] * max(1, keep.sum())
    assert np.linalg.norm(gv - np.eye(keep.sum())) <= 50 * TOL[dt] * max(1, keep.sum())


@pytest.mark.gpu
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,inds,inds_u", SVD_SHAPES)
def test_tensor_svd_thin_parity(shape, inds, inds_u, dt):
    import muscle_b200 as mb
    rng = np.random.default_rng(21)
    _check_svd(mb, random_array(rng, shape, dt), inds, inds_u, dt)


@pytest.mark.gpu
def test_tensor_svd_thin_rank_deficient_and_larger():
    import muscle_b200 as mb
    rng = np.random.default_rng(22)
    for dt in ("complex128", "float32"):
        low = random_array(rng, (48, 5), dt) @ random_array(rng, (5, 40), dt)     # rank 5
        _check_svd(mb, np.asfortranarray(low), "ab", "a", dt)
        _check_svd(mb, np.zeros((6, 4), dt), "ab", "a", dt)
    _check_svd(mb, random_array(rng, (256, 2, 192, 2), "complex128"), "lpqr", "lp", "complex128")   # 512 x 384
    assert mb.Handle.get(0).stats()["launches_svd"] > 0


@pytest.mark.gpu
@pytest.mark.parametrize("dt", DTYPES)
def test_tensor_svd_thin_is_isometric_on_rank_deficient_input(dt):
    """LAPACK (the reference's `svd`, tensor_svd.jl:113) returns orthonormal U and Vt whatever the rank; so must the device
    SVD: product states / low-entanglement bonds give exactly rank-deficient Theta, and canonical forms assume U'U = I."""
    import muscle_b200 as mb
    rng = np.random.default_rng(31)
    I = lambda s: [mb.Index(c) for c in s]
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    tol = 50 * TOL[dt]
    cases = [np.asfortranarray(random_array(rng, (40, 1), dt) @ random_array(rng, (1, 24), dt)),          # product state: rank 1
             np.asfortranarray(random_array(rng, (16, 3), dt) @ random_array(rng, (3, 48), dt)),          # wide, rank 3
             np.zeros((12, 12), dt),                                                                       # all-zero
             np.asfortranarray(np.diag(np.array([3, 0, 2, 0, 0, 1], dtype=dt)))]                           # exact zero columns
    for a in cases:
        U, S, Vt = mb.tensor_svd_thin(mb.Tensor(a, I("ab")).to_device(), inds_u=I("a"), ind_s=mb.Index("x"))
        u, s, vt = U.to_host().data.astype(wide), S.to_host().data, Vt.to_host().data.astype(wide)
        k = s.shape[0]
        info = mb.svd_last_info()
        assert info["converged"] and info["sweeps"] >= 0
        assert np.linalg.norm(u.conj().T @ u - np.eye(k)) <= tol * k, (a.shape, info)
        assert np.linalg.norm(vt.conj().T @ vt - np.eye(k)) <= tol * k, (a.shape, info)
        assert rel_frobenius((u * s.astype(wide)) @ vt.T, a.astype(wide)) <= TOL[dt] or not a.any()
        so = np.linalg.svd(a.astype(wide), compute_uv=False)
        assert np.linalg.norm(s - so) <= TOL[dt] * max(np.linalg.norm(so), 1e-300)
    assert mb.svd_last_info()["completed_columns"] >= 1


@pytest.mark.gpu
@pytest.mark.parametrize("dt,scale", [("float32", 1e-14), ("float32", 1e12), ("complex64", 3e-15), ("complex64", 2e11),
                                      ("float64", 1e-120), ("complex128", 1e130)])
def test_tensor_svd_thin_is_scale_invariant(dt, scale):
    """Squared norms and gamma^2 are 2nd / 4th powers of the data: without the power-of-two pre-scaling every rotation was
    skipped for Float32 data near 1e-11 (underflow) or 1e9 (overflow) and an un-diagonalised 'SVD' came back silently."""
    import muscle_b200 as mb
    rng = np.random.default_rng(32)
    I = lambda s: [mb.Index(c) for c in s]
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    a = np.asfortranarray((random_array(rng, (40, 28), dt).astype(wide) * scale).astype(dt))
    U, S, Vt = mb.tensor_svd_thin(mb.Tensor(a, I("ab")).to_device(), inds_u=I("a"), ind_s=mb.Index("x"))
    u, s, vt = U.to_host().data.astype(wide), S.to_host().data.astype(np.float64), Vt.to_host().data.astype(wide)
    so = np.linalg.svd(a.astype(wide), compute_uv=False)
    assert mb.svd_last_info()["converged"]
    assert np.linalg.norm(s - so) <= TOL[dt] * np.linalg.norm(so)
    assert rel_frobenius((u * s) @ vt.T, a.astype(wide)) <= TOL[dt]
    assert np.linalg.norm(u.conj().T @ u - np.eye(28)) <= 50 * TOL[dt] * 28


@pytest.mark.gpu
def test_tensor_svd_thin_reports_non_convergence():
    import muscle_b200 as mb
    rng = np.random.default_rng(33)
    I = lambda s: [mb.Index(c) for c in s]
    A = mb.Tensor(random_array(rng, (64, 64), "complex128"), I("ab")).to_device()
    mb.tensor_svd_thin(A, inds_u=I("a"), max_sweeps=1)
    info = mb.svd_last_info()
    assert info["sweeps"] == 1 and not info["converged"]
    mb.tensor_svd_thin(A, inds_u=I("a"))
    assert mb.svd_last_info()["converged"]


@pytest.mark.gpu
def test_simple_update_theta_needs_no_permute_pass():
    """simple_update asks the second contraction for Theta in [inds_u; inds_v] order (simple_update.jl:54-57), so no K1
    permutation runs between the contraction and the SVD."""
    import muscle_b200 as mb
    rng = np.random.default_rng(34)
    chi, d = 32, 2
    I = mb.Index
    a, b = random_array(rng, (chi, d, chi), "complex128"), random_array(rng, (chi, d, chi), "complex128")
    g = random_array(rng, (d, d, d, d), "complex128")
    A, B, G = (mb.Tensor(x, ix).to_device() for x, ix in ((a, [I("l"), I("pa"), I("bond")]), (b, [I("bond"), I("pb"), I("r")]),
                                                          (g, [I("pa"), I("pb"), I("ga"), I("gb")])))
    h = mb.Handle.get(0)
    h.reset_stats()
    mb.simple_update(A, I("pa"), B, I("pb"), I("bond"), G, I("ga"), I("gb"))
    st = h.stats()
    assert st["launches_permute"] == 0 and st["launches_svd"] == 1, st


@pytest.mark.gpu
def test_tensor_svd_thin_rejects_like_the_reference():
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    A = mb.Tensor(np.ones((2, 4, 6, 8)), I("ijkl")).to_device()
    for kw in (dict(), dict(inds_u=I("z")), dict(inds_v=I("z")), dict(inds_u=I("ijkl")), dict(inds_v=I("ijkl")),
               dict(inds_u=I("i"), ind_s=mb.Index("j"))):                           # tensor_svd_thin.jl:9-21
        with pytest.raises(mb.ArgumentError):
            mb.tensor_svd_thin(A, **kw)


@pytest.mark.gpu
def test_simple_update_reference_battery():
    """test/unit/operations/simple_update.jl:33-200 on BackendB200 (device tensors and host tensors via with_backend)."""
    import muscle_b200 as mb
    ga, gb, ident, cx = _gamma()
    Ia, Ib, Ibond, Iga, Igb = (mb.Index(("site", 1, "cut", 1)), mb.Index(("site", 2, "cut", 1)), mb.Index(("bond", 1, 2)),
                               mb.Index(("site", 1, "cut", 2)), mb.Index(("site", 2, "cut", 2)))
    for dev in (True, False):
        mk = (lambda x, ix: mb.Tensor(x, ix).to_device()) if dev else (lambda x, ix: mb.Tensor(x, ix))
        run = (lambda f: f()) if dev else (lambda f: mb.with_backend(f, mb.BackendB200()))
        Ga, Gb = mk(ga, [Ia, Ibond]), mk(gb, [Ib, Ibond])
        for gate, sv in ((ident, [1.0, 1.0]), (cx, [np.sqrt(2.0), 0.0])):
            Gt = mk(gate, [Ia, Ib, Iga, Igb])
            U, s, V = run(lambda: mb.simple_update(Ga, Ia, Gb, Ib, Ibond, Gt, Iga, Igb))
            assert U.inds == [Ia, Ibond] and V.inds == [Ib, Ibond] and s.inds == [Ibond]
            assert np.allclose(s.to_host().data, sv, atol=1e-14)
            psi = run(lambda: mb.binary_einsum(mb.hadamard(U, s), V))                 # :49-50
            assert np.isclose(np.sum(np.abs(psi.to_host().data) ** 2), 2.0)
            Un, sn, Vn = run(lambda: mb.simple_update(Ga, Ia, Gb, Ib, Ibond, Gt, Iga, Igb, normalize=True))
            assert np.isclose(np.linalg.norm(sn.to_host().data), 1.0)                 # normalize (:73-95)
            for absorb in (mb.AbsorbU(), mb.AbsorbV(), mb.AbsorbEqually()):           # absorb variants (:97-200)
                Ua, Va = run(lambda: mb.simple_update(Ga, Ia, Gb, Ib, Ibond, Gt, Iga, Igb, absorb=absorb))
                psi2 = run(lambda: mb.binary_einsum(Ua, Va))
                assert rel_frobenius(psi2.to_host().data, psi.to_host().data) <= 1e-13
            Um, sm, Vm = run(lambda: mb.simple_update(Ga, Ia, Gb, Ib, Ibond, Gt, Iga, Igb, maxdim=1))
            assert sm.shape == (1,) and Um.shape == (2, 1) and Vm.shape == (2, 1)
            assert np.isclose(sm.to_host().data[0], sv[0])


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["complex128", "complex64"])
def test_simple_update_random_mps_sites(dt):
    """Two random MPS sites (chi = 24, d = 2) and a random two-site gate: Θ and its singular values against the
    oracle (binary_einsum restatement + LAPACK), reconstruction of Θ from (U, s, V), truncation to maxdim."""
    import muscle_b200 as mb
    rng = np.random.default_rng(30)
    chi, d = 24, 2
    a, b = random_array(rng, (chi, d, chi), dt), random_array(rng, (chi, d, chi), dt)
    g = random_array(rng, (d, d, d, d), dt)
    I = mb.Index
    ia, ib, ig = [I("l"), I("pa"), I("bond")], [I("bond"), I("pb"), I("r")], [I("pa"), I("pb"), I("ga"), I("gb")]
    A, B, G = (mb.Tensor(x, ix).to_device() for x, ix in ((a, ia), (b, ib), (g, ig)))
    U, s, V = mb.simple_update(A, I("pa"), B, I("pb"), I("bond"), G, I("ga"), I("gb"))
    wide = np.complex128
    th, ith = simple_update_theta(a.astype(wide), ["l", "pa", "bond"], b.astype(wide), ["bond", "pb", "r"], g.astype(wide),
                                  ["pa", "pb", "ga", "gb"], "pa", "pb", "bond", "ga", "gb")
    _, so, _ = tensor_svd_thin_base(th, ith, ["l", "pa"], ["pb", "r"])
    tol = 1e-12 if dt == "complex128" else 2e-5
    assert U.inds == [I("l"), I("pa"), I("bond")] and V.inds == [I("pb"), I("r"), I("bond")]
    assert np.linalg.norm(s.to_host().data - so) <= tol * np.linalg.norm(so)
    rec = mb.binary_einsum(mb.hadamard(U, s), V, out=[I(c) for c in ith]).to_host().data
    assert rel_frobenius(rec.astype(wide), th) <= tol
    Ut, st, Vt = mb.simple_update(A, I("pa"), B, I("pb"), I("bond"), G, I("ga"), I("gb"), maxdim=chi)
    assert st.shape == (chi,) and Ut.shape == (chi, d, chi) and Vt.shape == (d, chi, chi)
    assert np.linalg.norm(st.to_host().data - so[:chi]) <= tol * np.linalg.norm(so)


# ---- tensor_qr_thin -------------------------------------------------------------------------------------------
def _check_qr(mb, a, inds, inds_q, dt):
    """test/unit/operations/tensor_qr_thin.jl:24-33: index / shape bookkeeping, Q·R ≈ A, Q isometric; plus R upper
    triangular. (QR is unique only up to a diagonal phase, so Q and R are not compared with LAPACK entry by entry.)"""
    I = lambda s: [mb.Index(c) for c in s]
    A = mb.Tensor(a, I(inds)).to_device()
    Q, R = mb.tensor_qr_thin(A, inds_q=I(inds_q), ind_virtual=mb.Index("x"))
    inds_r = [c for c in inds if c not in inds_q]
    left = tuple(a.shape[inds.index(c)] for c in inds_q)
    right = tuple(a.shape[inds.index(c)] for c in inds_r)
    k = min(int(np.prod(left)), int(np.prod(right)))
    assert Q.inds == I(inds_q) + [mb.Index("x")] and R.inds == [mb.Index("x")] + I(inds_r)
    assert Q.shape == left + (k,) and R.shape == (k,) + right
    wide = np.complex128 if np.dtype(dt).kind == "c" else np.float64
    q = Q.to_host().data.reshape(-1, k, order="F").astype(wide)
    r = R.to_host().data.reshape(k, -1, order="F").astype(wide)
    ref = np.transpose(a, [inds.index(c) for c in list(inds_q) + inds_r]).reshape(q.shape[0], r.shape[1], order="F").astype(wide)
    assert rel_frobenius(q @ r, ref) <= TOL[dt]
    assert np.linalg.norm(q.conj().T @ q - np.eye(k)) <= 50 * TOL[dt] * max(1, k)
    assert np.linalg.norm(np.tril(r, -1)) == 0.0
    # |R_jj| agrees with LAPACK's (the diagonal is fixed up to its phase)
    rl = np.linalg.qr(ref, mode="r")
    assert np.linalg.norm(np.abs(np.diag(r)) - np.abs(np.diag(rl))) <= 50 * TOL[dt] * np.linalg.norm(np.diag(rl))
    back = mb.binary_einsum(Q, R, out=I(inds)).to_host().data                   # :31 binary_einsum(Q, R) ≈ A
    assert rel_frobenius(back.astype(wide), a.astype(wide)) <= TOL[dt]


@pytest.mark.gpu
@pytest.mark.parametrize("dt", DTYPES)
@pytest.mark.parametrize("shape,inds,inds_q", SVD_SHAPES)
def test_tensor_qr_thin_parity(shape, inds, inds_q, dt):
    import muscle_b200 as mb
    rng = np.random.default_rng(41)
    _check_qr(mb, random_array(rng, shape, dt), inds, inds_q, dt)


@pytest.mark.gpu
def test_tensor_qr_thin_larger_and_rejects():
    import muscle_b200 as mb
    rng = np.random.default_rng(42)
    _check_qr(mb, random_array(rng, (256, 2, 160, 2), "complex128"), "lpqr", "lp", "complex128")    # 512 x 320
    _check_qr(mb, random_array(rng, (96, 700), "float32"), "ab", "a", "float32")
    I = lambda s: [mb.Index(c) for c in s]
    A = mb.Tensor(np.ones((2, 4, 6, 8)), I("ijkl")).to_device()
    for kw in (dict(), dict(inds_q=I("z")), dict(inds_r=I("z")), dict(inds_q=I("ijkl")), dict(inds_r=I("ijkl")),
               dict(inds_q=I("i"), ind_virtual=mb.Index("j"))):                 # tensor_qr_thin.jl:9-21
        with pytest.raises(mb.ArgumentError):
            mb.tensor_qr_thin(A, **kw)


@pytest.mark.gpu
def test_tensor_svd_trunc_matches_reference_rule():
    """tensor_svd.jl:153-201 / test/unit/operations/tensor_svd_trunc.jl: defaults = thin SVD; maxdim caps k; threshold
    cuts at the first singular value below threshold * norm(s) (that value is kept, as `keep = 1:findfirst(...)`)."""
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    rng = np.random.default_rng(50)
    a = random_array(rng, (200, 100), "complex128")
    A = mb.Tensor(a, I("ij")).to_device()
    U, s, Vt = mb.tensor_svd_trunc(A, inds_u=I("i"), ind_s=mb.Index("x"))
    assert s.shape == (100,) and U.shape == (200, 100) and Vt.shape == (100, 100)
    rec = mb.binary_einsum(mb.hadamard(U, s), Vt, out=I("ij")).to_host().data
    assert rel_frobenius(rec, a) <= 1e-12                                        # tensor_svd_trunc.jl:38-40
    so = np.linalg.svd(a, compute_uv=False)
    U, s, Vt = mb.tensor_svd_trunc(A, inds_u=I("i"), ind_s=mb.Index("x"), maxdim=17)
    assert s.shape == (17,) and U.shape == (200, 17) and Vt.shape == (100, 17)
    assert np.linalg.norm(s.to_host().data - so[:17]) <= 1e-12 * np.linalg.norm(so)
    thr = 0.08
    k_ref = int(np.nonzero(so < np.linalg.norm(so) * thr)[0][0]) + 1
    U, s, Vt = mb.tensor_svd_trunc(A, inds_u=I("i"), ind_s=mb.Index("x"), threshold=thr)
    assert s.shape == (k_ref,) and U.shape == (200, k_ref)
    # the truncated factors are the leading part of the full ones: best rank-k approximation error = tail norm
    rec = mb.binary_einsum(mb.hadamard(U, s), Vt, out=I("ij")).to_host().data
    assert abs(np.linalg.norm(rec - a) - np.linalg.norm(so[k_ref:])) <= 1e-10 * np.linalg.norm(so)
