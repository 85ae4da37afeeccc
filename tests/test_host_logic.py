"""CPU: the C-ABI library loads and exports every symbol of include/muscle_b200.h, the planner's
index bookkeeping, the front-end kwargs semantics and the Backend/Domain dispatch mirror.
No compute call is made here (there is no GPU and no CPU compute path in the product)."""
import os
import re

import numpy as np
import pytest

import muscle_b200 as mb
from muscle_b200 import (ArgumentError, B200Error, BackendB200, BackendBase, DimensionMismatch, Index, Tensor,
                         binary_einsum, binary_einsum_, with_backend)
from muscle_b200 import _lib
from cases import PARITY_CASES, REFERENCE_BATTERY

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_header_symbol():
    header = open(os.path.join(ROOT, "include", "muscle_b200.h")).read()
    declared = set(re.findall(r"\b(mb200_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 20
    L = mb.lib()
    for name in declared:
        assert hasattr(L, name), f"{name} declared in include/muscle_b200.h but not exported"
    assert declared == set(_lib.PROTOTYPES), "ctypes prototypes out of sync with the header"
    assert L.mb200_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    A = Tensor(np.ones((2, 3)), "ij")
    B = Tensor(np.ones((3, 4)), "jk")
    with pytest.raises(B200Error):
        binary_einsum(BackendB200(), [Index("i"), Index("k")], A, B)


def test_frontend_matches_reference_battery():
    """kwargs → inds_c (binary_einsum.jl:33-41), replaying the reference's test inputs."""
    for name, sa, ia, sb, ib, kw, exp_inds, exp_shape, _, _ in REFERENCE_BATTERY:
        kwargs = {k: [Index(c) for c in v] for k, v in kw.items()}
        got = mb.frontend_inds_c([Index(c) for c in ia], [Index(c) for c in ib], **kwargs)
        assert got == [Index(c) for c in exp_inds], name


def test_frontend_matches_oracle_frontend():
    from oracle import frontend_inds_c as oracle_frontend
    rng = np.random.default_rng(0)
    labels = list("abcdefgh")
    for _ in range(200):
        ia = list(rng.permutation(labels)[: rng.integers(0, 6)])
        ib = list(rng.permutation(labels)[: rng.integers(0, 6)])
        shared = [x for x in ia if x in ib]
        dims = None if rng.random() < 0.3 else list(rng.permutation(labels)[: rng.integers(0, 4)])
        exp = oracle_frontend(ia, ib, dims=dims)
        got = mb.frontend_inds_c([Index(c) for c in ia], [Index(c) for c in ib],
                                 dims=None if dims is None else [Index(c) for c in dims])
        assert [g.tag for g in got] == exp, (ia, ib, dims, shared)


def test_flatten_labels_like_cutensor_ext():
    """indmap = enumerate(unique(inds_a ∪ inds_b)) — ext/MuscleCUDAExt.jl:24-27."""
    ia = [Index(c) for c in "ijk"]
    ib = [Index(c) for c in "klj"]
    ma, mb_, mc = mb.flatten_labels(ia, ib, [Index("i"), Index("l")])
    assert (ma, mb_, mc) == ([0, 1, 2], [2, 3, 1], [0, 3])
    with pytest.raises(ArgumentError):
        mb.flatten_labels(ia, ib, [Index("z")])
    # tags may be any hashable (symbols, ints, named tuples — test/unit/operations/simple_update.jl:8)
    assert Index(("site", 1)) == Index(("site", 1)) and Index(1) != Index("1")


def test_backend_dispatch():
    A = Tensor(np.ones((2, 3)), "ij")
    B = Tensor(np.ones((3, 4)), "jk")
    assert mb.choose_backend("binary_einsum", A.parent, B.parent) == BackendBase()   # binary_einsum.jl:20
    assert with_backend(lambda: mb.choose_backend("binary_einsum", A.parent, B.parent), BackendB200()) == BackendB200()
    # scoped: the override ends with the call (ScopedValue, src/Backend.jl:16-18)
    assert mb.choose_backend("binary_einsum", A.parent, B.parent) == BackendBase()
    # a backend without a method → ArgumentError (binary_einsum.jl:53-55,72-74)
    with pytest.raises(ArgumentError):
        binary_einsum(A, B)
    with pytest.raises(ArgumentError):
        binary_einsum_(Tensor(np.zeros((2, 4)), "ik"), A, B)
    assert mb.domain(A) == mb.DomainHost()


def test_tensor_ctor_checks():
    with pytest.raises(ArgumentError):
        Tensor(np.ones((2, 3)), "i")
    with pytest.raises(DimensionMismatch):
        Tensor(np.ones((2, 3)), "ii")
    t = Tensor(np.ones((2, 3, 4)), "ijk")
    assert t.inds == [Index("i"), Index("j"), Index("k")] and t.size(Index("k")) == 4 and t.dim("j") == 1
    p = t.permutedims([Index("k"), Index("i"), Index("j")])
    assert p.shape == (4, 2, 3) and p.inds == [Index("k"), Index("i"), Index("j")]
    assert t.isequal(p) and t.isapprox(p)


def _describe(ext, ia, ib, ic, dt=_lib.C128):
    labels = {c: k for k, c in enumerate(dict.fromkeys(ia + ib))}
    return mb.plan_describe(dt, [labels[c] for c in ic], dt, [labels[c] for c in ia], [ext[c] for c in ia],
                            dt, [labels[c] for c in ib], [ext[c] for c in ib]), labels


@pytest.mark.parametrize("case", PARITY_CASES, ids=[c[0] for c in PARITY_CASES])
def test_planner_classification(case):
    name, ext, ia, ib, ic = case
    info, labels = _describe(ext, ia, ib, ic)
    inv = {v: k for k, v in labels.items()}
    left = [inv[info.left[i]] for i in range(info.n_left)]
    right = [inv[info.right[i]] for i in range(info.n_right)]
    summed = [inv[info.sum[i]] for i in range(info.n_sum)]
    batch = [inv[info.batch[i]] for i in range(info.n_batch)]
    row, col = (ib, ia) if info.swapped else (ia, ib)
    live = lambda s: {c for c in s if ext[c] > 1}
    assert set(batch) == live(set(ia) & set(ib) & set(ic))
    assert set(summed) == live((set(ia) & set(ib)) - set(ic))
    assert set(left) == live(set(row) - set(col))
    assert set(right) == live(set(col) - set(row))
    prod = lambda s: int(np.prod([ext[c] for c in s], dtype=np.int64)) if s else 1
    assert (info.M, info.N, info.K, info.L) == (prod(left), prod(right), prod(summed), prod(batch))
    assert info.flops == 8.0 * info.M * info.N * info.K * info.L
    # C's unit-stride live mode is always a row (or batch) mode: that is what `swapped` guarantees
    live_c = [c for c in ic if ext[c] > 1]
    if live_c:
        assert live_c[0] in left + batch
    # walk order of the row/column/batch groups follows C's memory order
    for grp in (left, right, batch):
        pos = [ic.index(c) for c in grp]
        assert pos == sorted(pos)


def test_planner_sizes_of_baseline_configs():
    """GEMM-equivalent sizes and paths of the north-star configs."""
    cfg1, _ = _describe(dict(i=64, j=64, k=64, l=64, m=64, n=64), "kilj", "nlmk", "mjni")
    assert (cfg1.M, cfg1.N, cfg1.K, cfg1.L) == (4096, 4096, 4096, 1) and cfg1.path == mb.PATH_GETT_F64
    c2a, _ = _describe(dict(a=1024, w=8, b=1024, s=2, c=1024), "awb", "bsc", "awsc")
    assert (c2a.M, c2a.N, c2a.K) == (8192, 2048, 1024) and c2a.flops == 8.0 * 8192 * 2048 * 1024
    c2b, _ = _describe(dict(a=1024, w=8, s=2, c=1024, t=2, v=8), "awsc", "wstv", "atvc")
    assert (c2b.M, c2b.N, c2b.K) == (1024 * 1024, 16, 16)
    c2c, _ = _describe(dict(a=1024, t=2, v=8, c=1024, e=1024), "atvc", "ate", "evc")
    assert sorted((c2c.M, c2c.N)) == [1024, 8192] and c2c.K == 2048
    c3, _ = _describe(dict(l=256, k=8, b=8, m=256, q=8, r=256, z=8), "lkbmz", "mkqrz", "lbqrz", dt=_lib.C64)
    assert (c3.M, c3.N, c3.K, c3.L) == (2048, 2048, 2048, 8) and c3.path in (mb.PATH_SIMT_F32, mb.PATH_TCGEN05_TF32)
    tiny, _ = _describe(dict(i=2, j=3, k=4), "ij", "jk", "ik")
    assert tiny.path == mb.PATH_DIRECT   # everything in the reference's own test-suite is launch-bound


def test_planner_rejections():
    d = _lib.F64
    with pytest.raises(ArgumentError):      # repeated label inside one operand (ext/MuscleReactantExt.jl:88-89)
        mb.plan_describe(d, [0], d, [0, 0], [2, 2], d, [1], [3])
    with pytest.raises(ArgumentError):      # label of C in neither operand (ext/MuscleStridedExt.jl:53)
        mb.plan_describe(d, [0, 5], d, [0, 1], [2, 3], d, [1], [3])
    # a free label missing from C: BackendBase rejects it (binary_einsum.jl:83); BackendB200 sums it like cuTENSOR / OMEinsum
    # (ext/MuscleCUDAExt.jl:30-38) - the plan is the contraction left after the pre-reduce
    pre = mb.plan_describe(d, [0], d, [0, 1], [2, 3], d, [1, 2], [3, 4])
    assert (pre.M, pre.N, pre.K) == (2, 1, 3)
    with pytest.raises(DimensionMismatch):  # shared label with different extents
        mb.plan_describe(d, [0], d, [0, 1], [2, 3], d, [1], [4])
    with pytest.raises(ArgumentError):      # eltype of C must be the promotion
        mb.plan_describe(_lib.F32, [0], d, [0, 1], [2, 3], d, [1], [3])
    with pytest.raises(ArgumentError):      # repeated label in C
        mb.plan_describe(d, [0, 0], d, [0, 1], [2, 3], d, [1], [3])
    ok = mb.plan_describe(_lib.C128, [0], _lib.F64, [0, 1], [2, 3], _lib.C128, [1], [3])
    assert ok.compute_dtype == _lib.C128     # mixed eltypes promote (src/Muscle.jl:47 precompile workload)


def test_errors_surface_before_any_device_work():
    """Argument errors must not depend on a GPU: they are raised by the host-only planner."""
    A = Tensor(np.ones((2, 3)), "ij")
    B = Tensor(np.ones((4, 5)), "jk")
    with pytest.raises(DimensionMismatch):
        binary_einsum(BackendB200(), [Index("i"), Index("k")], A, B)
    with pytest.raises(ArgumentError):      # a label of the output found in neither operand
        binary_einsum(BackendB200(), [Index("i"), Index("q")], A, Tensor(np.ones((3, 5)), "jk"))
    with pytest.raises(ArgumentError):
        binary_einsum(BackendB200(), [Index("i"), Index("k")], Tensor(np.ones((2, 3), np.int64), "ij"),
                      Tensor(np.ones((3, 5)), "jk"))


def test_shard_plan():
    # cfg 4: free-index shard, no collective
    labels = {c: k for k, c in enumerate("abcdefghi")}
    ia, ib, ic = "adbecf", "fgdhei", "abcghi"
    ext = {c: 16 for c in labels}
    m = lambda s: [labels[c] for c in s]
    e = lambda s: [ext[c] for c in s]
    covered = []
    for r in range(8):
        s = mb.shard_plan(m(ic), m(ia), e(ia), m(ib), e(ib), 8, r)
        assert s.kind == _lib.SHARD_FREE and s.mode == labels["i"] and not s.needs_allreduce
        covered.append((s.begin, s.end))
    assert covered == [(2 * r, 2 * r + 2) for r in range(8)]
    # cfg 5: summed-index slice + allreduce
    s = mb.shard_plan(m(ic), m(ia), e(ia), m(ib), e(ib), 8, 3, prefer_sum=True)
    assert s.kind == _lib.SHARD_SUM and s.needs_allreduce and s.mode == labels["f"] and (s.begin, s.end) == (6, 8)
    # batch mode
    s = mb.shard_plan([0, 2, 3], [0, 1, 3], [4, 5, 8], [1, 2, 3], [5, 6, 8], 4, 1)
    assert s.kind == _lib.SHARD_BATCH and s.mode == 3 and (s.begin, s.end) == (2, 4)
    # uneven split covers the whole range without overlap
    spans = [mb.shard_plan([0, 2], [0, 1], [7, 3], [1, 2], [3, 10], 4, r) for r in range(4)]
    assert spans[0].begin == 0 and spans[-1].end == 10
    assert all(spans[i].end == spans[i + 1].begin for i in range(3))
    # nothing shardable → replicas only
    s = mb.shard_plan([0], [0, 1], [2, 3], [1], [3], 8, 0)
    assert s.kind == _lib.SHARD_NONE


def test_planner_tcgen05_eligibility_and_pack_kind():
    """ComplexF32 / Float32 shapes with M >= 64, N >= 32, K >= 64 are tcgen05-eligible whatever the summed extents and strides;
    the operands are packed by a K1 permutation when the leading summed modes tile groups of 8 k and both tensors are dense,
    by the table-driven gather pack otherwise (host-only planner facts, no device needed)."""
    C64, F32, C128 = _lib.C64, _lib.F32, _lib.C128
    # config 3 (PEPS, D = 8): permutation pack
    i = mb.plan_describe(C64, [0, 2, 4, 5, 6], C64, [0, 1, 2, 3, 6], [256, 8, 8, 256, 8], C64, [3, 1, 4, 5, 6], [256, 8, 8, 256, 8])
    assert i.path == mb.PATH_TCGEN05_TF32 and i.tc_eligible == 1 and i.tc_permute_pack == 1
    # K = 100 (4 | 100 but 8 does not): eligible, gather pack; big enough -> tcgen05 picked
    i = mb.plan_describe(C64, [1, 2], C64, [0, 1], [100, 2048], C64, [0, 2], [100, 2048])
    assert i.tc_eligible == 1 and i.tc_permute_pack == 0 and i.path == mb.PATH_TCGEN05_TF32
    # odd bond dimensions 3 x 5 x 7 = 105
    i = mb.plan_describe(F32, [3, 4], F32, [0, 3, 1, 2], [3, 1300, 5, 7], F32, [2, 4, 1, 0], [7, 1100, 5, 3])
    assert i.K == 105 and i.tc_eligible == 1 and i.tc_permute_pack == 0 and i.path == mb.PATH_TCGEN05_TF32
    # a strided operand (every other row of a wider buffer): gather pack
    i = mb.plan_describe(C64, [1, 2], C64, [0, 1], [512, 2048], C64, [0, 2], [512, 1024], strides_b=[2, 1024])
    assert i.tc_eligible == 1 and i.tc_permute_pack == 0
    i = mb.plan_describe(C64, [1, 2], C64, [0, 1], [512, 2048], C64, [0, 2], [512, 1024])
    assert i.tc_eligible == 1 and i.tc_permute_pack == 1
    # K < 64, a skinny side, or a double-precision dtype: not eligible
    assert mb.plan_describe(C64, [1, 2], C64, [0, 1], [60, 2048], C64, [0, 2], [60, 2048]).tc_eligible == 0
    assert mb.plan_describe(C64, [1, 2], C64, [0, 1], [512, 2048], C64, [0, 2], [512, 16]).tc_eligible == 0
    assert mb.plan_describe(C128, [1, 2], C128, [0, 1], [512, 2048], C128, [0, 2], [512, 2048]).tc_eligible == 0


def test_bench_headline_shard_and_oracle_helpers():
    """bench.py's headline (config 4b, 16384^3) asks mb200_shard_plan for the cut at every N and asserts it is the free label `i`;
    its CPU-side parity helper must agree with the general oracle when a batch label is restricted to one value."""
    import importlib.util
    import os
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(os.path.dirname(os.path.dirname(__file__)), "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    ext, ia, ib, ic = bench.CFG4B["ext"], bench.CFG4B["ia"], bench.CFG4B["ib"], bench.CFG4B["ic"]
    lab = "abcdefghi"
    for world in (2, 4, 8):
        cover = []
        for rank in range(world):
            info = _lib.shard_plan([lab.index(x) for x in ic], [lab.index(x) for x in ia], [ext[x] for x in ia],
                                   [lab.index(x) for x in ib], [ext[x] for x in ib], world, rank)
            assert info.kind == _lib.SHARD_FREE and lab[info.mode] == "i" and not info.needs_allreduce
            cover += list(range(info.begin, info.end))
        assert cover == list(range(ext["i"]))            # the slabs tile the label exactly once
    from oracle import binary_einsum_general
    rng = np.random.default_rng(0)
    a = rng.standard_normal((3, 4, 1)) + 1j * rng.standard_normal((3, 4, 1))
    b = rng.standard_normal((4, 5, 1)) + 1j * rng.standard_normal((4, 5, 1))
    ref, ic2 = bench.oracle_slab("ijz", "jkz", "kiz", np.asfortranarray(a), np.asfortranarray(b))
    full = binary_einsum_general(list("kiz"), a, list("ijz"), b, list("jkz"))
    assert ic2 == ["k", "i"] and np.allclose(ref, full[..., 0])
    A, B, e2, flops = bench.cfg4b_sample(1 / 16, 1 / 16)
    assert A.shape == tuple(e2[x] for x in ia) and e2["a"] == 2 and e2["g"] == 2 and flops == 8.0 * np.prod([e2[x] for x in e2])


def test_reference_arm_uses_every_host_thread_whatever_the_launcher_exported():
    """torchrun exports OMP_NUM_THREADS=1; round 1's N > 1 reference arm therefore ran single-threaded. bench.set_blas_threads() asks
    threadpoolctl for os.cpu_count() BLAS threads regardless."""
    import importlib.util
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = ("import importlib.util,os;s=importlib.util.spec_from_file_location('b',r'%s');b=importlib.util.module_from_spec(s);"
            "s.loader.exec_module(b);import numpy;print(b.set_blas_threads(), os.cpu_count())" % os.path.join(root, "bench.py"))
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, env=dict(os.environ, OMP_NUM_THREADS="1"), timeout=120)
    got, ncpu = map(int, out.stdout.split()[-2:])
    assert got == ncpu or ncpu == 1, out.stdout + out.stderr


def test_planner_classification_fuzz():
    """Randomised label bookkeeping (hypothesis): for arbitrary label sets, extents (incl. 1), operand orders and output orders - with
    dangling labels thrown in - the planner's batch / summed / free classification, GEMM sizes, swap rule and walk order agree with the
    set arithmetic of the reference's front-end and backends (binary_einsum.jl:33-41, 76-79; ext/MuscleReactantExt.jl:94-109)."""
    from hypothesis import given, settings, strategies as st

    letters = "abcdefghijkl"

    @st.composite
    def problem(draw):
        n = draw(st.integers(2, 9))
        labs = list(letters[:n])
        ext = {c: draw(st.sampled_from([1, 2, 3, 4, 5, 8])) for c in labs}
        # where each label lives: 1 = A only, 2 = B only, 3 = both
        where = {c: draw(st.sampled_from([1, 2, 3])) for c in labs}
        ia = draw(st.permutations([c for c in labs if where[c] & 1]))
        ib = draw(st.permutations([c for c in labs if where[c] & 2]))
        # output: every label may be kept or dropped (dropping a single-operand label makes it dangling: summed by the entry points)
        keep = [c for c in labs if draw(st.booleans())]
        ic = draw(st.permutations(keep))
        return ext, "".join(ia), "".join(ib), "".join(ic)

    @settings(max_examples=300, deadline=None)
    @given(problem())
    def check(pr):
        ext, ia, ib, ic = pr
        info, labels = _describe(ext, ia, ib, ic)
        inv = {v: k for k, v in labels.items()}
        left = [inv[info.left[i]] for i in range(info.n_left)]
        right = [inv[info.right[i]] for i in range(info.n_right)]
        summed = [inv[info.sum[i]] for i in range(info.n_sum)]
        batch = [inv[info.batch[i]] for i in range(info.n_batch)]
        live = lambda s: {c for c in s if ext[c] > 1}
        sa, sb, sc = set(ia), set(ib), set(ic)
        dangling = ((sa - sb) | (sb - sa)) - sc           # summed away before planning: in none of the plan's groups
        row, col = (ib, ia) if info.swapped else (ia, ib)
        assert set(batch) == live(sa & sb & sc)
        assert set(summed) == live((sa & sb) - sc)
        assert set(left) == live(set(row) - set(col) - dangling)
        assert set(right) == live(set(col) - set(row) - dangling)
        prod = lambda s: int(np.prod([ext[c] for c in s], dtype=np.int64)) if s else 1
        assert (info.M, info.N, info.K, info.L) == (prod(left), prod(right), prod(summed), prod(batch))
        live_c = [c for c in ic if ext[c] > 1]
        if live_c:
            assert live_c[0] in left + batch              # C's unit-stride live label is a row (or batch) label
        for grp in (left, right, batch):
            pos = [ic.index(c) for c in grp]
            assert pos == sorted(pos)                     # groups are walked in C's memory order

    check()


def test_shard_plan_fuzz():
    """mb200_shard_plan (Dagger's block scheme, ext/MuscleDaggerExt/binary_einsum.jl:64-119): for random contractions and rank counts the
    per-rank ranges tile the chosen label exactly once, every rank picks the same label and kind, a summed cut asks for the add-reduction
    and a free / batch cut does not, and prefer_sum only ever picks a label that is summed."""
    from hypothesis import given, settings, strategies as st

    @st.composite
    def problem(draw):
        n = draw(st.integers(2, 8))
        labs = list("abcdefgh"[:n])
        ext = {c: draw(st.sampled_from([1, 2, 3, 4, 6, 8, 16])) for c in labs}
        where = {c: draw(st.sampled_from([1, 2, 3])) for c in labs}
        ia = [c for c in labs if where[c] & 1]
        ib = [c for c in labs if where[c] & 2]
        ic = [c for c in labs if where[c] != 3 or draw(st.booleans())]     # single-operand labels are kept, shared ones kept (batch) or summed
        return ext, ia, ib, draw(st.permutations(ic)), draw(st.integers(1, 8)), draw(st.booleans())

    @settings(max_examples=200, deadline=None)
    @given(problem())
    def check(pr):
        ext, ia, ib, ic, nranks, prefer_sum = pr
        ids = {c: k for k, c in enumerate("abcdefgh")}
        infos = [_lib.shard_plan([ids[c] for c in ic], [ids[c] for c in ia], [ext[c] for c in ia], [ids[c] for c in ib],
                                 [ext[c] for c in ib], nranks, r, prefer_sum) for r in range(nranks)]
        assert len({(i.kind, i.mode) for i in infos}) == 1
        kind, mode = infos[0].kind, infos[0].mode
        if kind == _lib.SHARD_NONE:
            return
        lab = "abcdefgh"[mode]
        cover = [x for i in infos for x in range(i.begin, i.end)]
        assert cover == list(range(ext[lab])) and ext[lab] >= nranks
        in_a, in_b, in_c = lab in ia, lab in ib, lab in ic
        if kind == _lib.SHARD_SUM:
            assert in_a and in_b and not in_c and all(i.needs_allreduce for i in infos)
        elif kind == _lib.SHARD_BATCH:
            assert in_a and in_b and in_c and not any(i.needs_allreduce for i in infos)
        else:
            assert (in_a != in_b) and in_c and not any(i.needs_allreduce for i in infos)

    check()
