"""Blocked tensors, the Dagger bridge: `BackendBlocks` mirrors `Dagger.stage(::BinaryEinsum)`
(ext/MuscleDaggerExt/binary_einsum.jl:64-119). CPU tests: the oracle restatement pinned on the reference's own test
(test/integration/dagger.jl:12-31), the block bookkeeping with the chunk contraction injected, dispatch and the error
cases. GPU tests: the same cases with device chunks through BackendB200 (every chunk contraction and the slot-sum
reduction run on the B200)."""
import numpy as np
import pytest

import muscle_b200 as mb
from muscle_b200 import BackendB200, BackendBlocks, BlockArray, Blocks, Index, Tensor, binary_einsum, distribute, with_backend
from muscle_b200 import einsum as einsum_mod
from cases import random_array
from oracle import binary_einsum_general, dagger_stage_oracle, rel_frobenius

I = lambda s: [Index(c) for c in s]

# (name, extents, ia, blocks_a, ib, blocks_b, ic)
BLOCK_CASES = [
    ("dagger_jl_block_block", dict(i=2, j=2, k=2), "ij", (1, 1), "jk", (1, 1), "ik"),        # test/integration/dagger.jl:12-31
    ("matmul_2x3_sum_blocks", dict(i=8, j=12, k=6), "ij", (4, 4), "jk", (4, 3), "ik"),
    ("out_transposed", dict(i=8, j=12, k=6), "ij", (4, 4), "jk", (4, 3), "ki"),
    ("rank3_two_summed", dict(a=4, b=6, c=4, d=6, e=2), "abc", (2, 3, 2), "cbde", (2, 3, 3, 2), "aed"),
    ("batch_label", dict(i=6, j=8, k=4, z=4), "ijz", (3, 4, 2), "jkz", (4, 2, 2), "kiz"),
    ("outer_product", dict(i=4, j=6), "i", (2,), "j", (3,), "ji"),
    ("single_block", dict(i=5, j=7, k=3), "ij", (5, 7), "jk", (7, 3), "ik"),
]


def _case(case, dt, seed=3):
    name, ext, ia, ba, ib, bb, ic = case
    rng = np.random.default_rng(seed)
    a = random_array(rng, tuple(ext[c] for c in ia), dt)
    b = random_array(rng, tuple(ext[c] for c in ib), dt)
    return a, list(ia), ba, b, list(ib), bb, list(ic)


def test_oracle_pinned_on_reference_dagger_test():
    """test/integration/dagger.jl:12-31 verbatim: 2 x 2 Float64 data in 1 x 1 blocks; `collect(block_c) ≈ c`, chunks (1, 1)."""
    data1 = np.array([[1.0, 2.0], [3.0, 4.0]])
    data2 = np.array([[5.0, 6.0], [7.0, 8.0]])
    c, bs, shapes = dagger_stage_oracle("ik", data1, "ij", (1, 1), data2, "jk", (1, 1))
    assert np.allclose(c, data1 @ data2) and bs == (1, 1) and all(s == (1, 1) for s in shapes) and len(shapes) == 4


@pytest.mark.parametrize("dt", ["float64", "complex128"])
@pytest.mark.parametrize("case", BLOCK_CASES, ids=[c[0] for c in BLOCK_CASES])
def test_oracle_stage_equals_dense_einsum(case, dt):
    a, ia, ba, b, ib, bb, ic = _case(case, dt)
    c, bs, shapes = dagger_stage_oracle(ic, a, ia, ba, b, ib, bb)
    assert rel_frobenius(c, binary_einsum_general(ic, a, ia, b, ib)) <= 1e-13
    assert all(s == bs for s in shapes)


def test_distribute_collect_and_dispatch():
    x = np.arange(24.0).reshape(4, 6)
    bx = distribute(x, Blocks(2, 3))
    assert isinstance(bx, BlockArray) and bx.shape == (4, 6) and bx.chunks.shape == (2, 2) and not bx.on_device
    assert bx.domainchunks() == [(2, 3)] * 4 and np.array_equal(bx.collect(), x)
    t = Tensor(bx, I("ij"))
    assert t.shape == (4, 6) and isinstance(mb.domain(t), mb.DomainBlocks)
    # Dagger rules, also mixed with a plain array (src/Operations/binary_einsum.jl:25-31)
    assert isinstance(mb.choose_backend("binary_einsum", bx, bx), BackendBlocks)
    assert isinstance(mb.choose_backend("binary_einsum", bx, x), BackendBlocks)
    assert isinstance(mb.choose_backend("binary_einsum", x, bx), BackendBlocks)
    with pytest.raises(mb.ArgumentError):
        distribute(x, Blocks(3, 3))            # 4 is not a multiple of 3 (Dagger.stage divides with ÷)
    with pytest.raises(mb.ArgumentError):
        distribute(x, Blocks(2))


def _inject_oracle(monkeypatch):
    """The chunk contraction needs a B200; on a CPU box the oracle stands in for it (as tests/test_dist_gloo.py does)."""
    def fake(inds_c, a, b):
        tags = lambda t: [i.tag for i in t.inds]
        c = binary_einsum_general([i.tag for i in inds_c], np.asarray(a.data), tags(a), np.asarray(b.data), tags(b))
        return Tensor(c, inds_c)
    monkeypatch.setattr(einsum_mod, "_b200_out_of_place", fake)


@pytest.mark.parametrize("dt", ["float64", "complex64"])
@pytest.mark.parametrize("case", BLOCK_CASES, ids=[c[0] for c in BLOCK_CASES])
def test_block_bookkeeping_with_injected_contraction(case, dt, monkeypatch):
    """Grid, block sizes, chunk selection over output / summed blocks and the host add tree, chunk contraction injected."""
    _inject_oracle(monkeypatch)
    a, ia, ba, b, ib, bb, ic = _case(case, dt)
    ta, tb = Tensor(distribute(a, Blocks(*ba)), I(ia)), Tensor(distribute(b, Blocks(*bb)), I(ib))
    tc = with_backend(lambda: binary_einsum(BackendBlocks(), I(ic), ta, tb), BackendB200())   # host chunks need the override
    ref, bs, shapes = dagger_stage_oracle(ic, a, ia, ba, b, ib, bb)
    assert isinstance(tc.parent, BlockArray) and tc.inds == I(ic)
    assert tc.parent.blocksize == bs and tc.parent.domainchunks() == shapes
    assert rel_frobenius(tc.parent.collect(), ref) <= (1e-13 if dt == "float64" else 1e-5)


def test_blocked_errors(monkeypatch):
    _inject_oracle(monkeypatch)
    a = distribute(np.ones((4, 6)), Blocks(2, 3))
    b = distribute(np.ones((6, 4)), Blocks(2, 2))          # summed label j: blocks of 3 against blocks of 2
    with pytest.raises(mb.ArgumentError):
        binary_einsum(BackendBlocks(), I("ik"), Tensor(a, I("ij")), Tensor(b, I("jk")))
    b2 = distribute(np.ones((6, 4)), Blocks(3, 2))
    with pytest.raises(mb.ArgumentError):                  # ic ⊄ ia ∪ ib (binary_einsum.jl:22)
        binary_einsum(BackendBlocks(), I("iq"), Tensor(a, I("ij")), Tensor(b2, I("jk")))
    with pytest.raises(mb.ArgumentError):                  # repeated output label (binary_einsum.jl:21)
        binary_einsum(BackendBlocks(), I("ii"), Tensor(a, I("ij")), Tensor(b2, I("jk")))
    with pytest.raises(mb.ArgumentError):                  # free label missing from the output
        binary_einsum(BackendBlocks(), I("i"), Tensor(a, I("ij")), Tensor(b2, I("jk")))


# ---- on the B200: device chunks, every chunk contraction on BackendB200, summed blocks reduced by mb200_reduce_slots ----
@pytest.mark.gpu
def test_reference_dagger_test_on_device():
    """test/integration/dagger.jl:12-31 with device chunks: `parent(block_c) isa DArray`, chunks (1, 1), `collect ≈ c`."""
    data1 = np.array([[1.0, 2.0], [3.0, 4.0]])
    data2 = np.array([[5.0, 6.0], [7.0, 8.0]])
    a, b = Tensor(data1, I("ij")).to_device(), Tensor(data2, I("jk")).to_device()
    block_a = Tensor(distribute(data1, Blocks(1, 1), device=0), I("ij"))
    block_b = Tensor(distribute(data2, Blocks(1, 1), device=0), I("jk"))
    c = binary_einsum(a, b)
    block_c = binary_einsum(block_a, block_b)              # Domain rule selects BackendBlocks
    assert isinstance(block_c.parent, BlockArray) and block_c.parent.on_device
    assert all(s == (1, 1) for s in block_c.parent.domainchunks())
    assert np.allclose(block_c.parent.collect(), c.to_host().data)


@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["float64", "complex128", "complex64", "float32"])
@pytest.mark.parametrize("case", BLOCK_CASES, ids=[c[0] for c in BLOCK_CASES])
def test_blocked_parity_on_device(case, dt):
    a, ia, ba, b, ib, bb, ic = _case(case, dt)
    ta = Tensor(distribute(a, Blocks(*ba), device=0), I(ia))
    tb = Tensor(distribute(b, Blocks(*bb), device=0), I(ib))
    h = mb.Handle.get(0)
    h.reset_stats()
    tc = binary_einsum(ta, tb, out=I(ic))
    ref, bs, shapes = dagger_stage_oracle(ic, a, ia, ba, b, ib, bb)
    assert tc.parent.on_device and tc.parent.blocksize == bs and tc.parent.domainchunks() == shapes
    assert h.stats()["launches_total"] > 0
    tol = 1e-12 if dt in ("float64", "complex128") else 1e-5
    assert rel_frobenius(tc.parent.collect(), ref) <= tol


@pytest.mark.gpu
def test_blocked_mixed_with_dense_and_large_chunks():
    """A blocked operand against a plain device array (one block), chunks large enough for the DMMA gather-GEMM."""
    rng = np.random.default_rng(6)
    a = random_array(rng, (256, 192), "complex128")
    b = random_array(rng, (192, 160), "complex128")
    ta = Tensor(distribute(a, Blocks(128, 192), device=0), I("ij"))
    tb = Tensor(b, I("jk")).to_device()
    tc = binary_einsum(ta, tb)
    assert tc.parent.chunks.shape == (2, 1) and rel_frobenius(tc.parent.collect(), a @ b) <= 1e-12
    # several summed blocks: the slot-sum reduction
    ta2 = Tensor(distribute(a, Blocks(128, 64), device=0), I("ij"))
    tb2 = Tensor(distribute(b, Blocks(64, 80), device=0), I("jk"))
    tc2 = binary_einsum(ta2, tb2, out=I("ki"))
    assert tc2.parent.chunks.shape == (2, 2) and rel_frobenius(tc2.parent.collect(), (a @ b).T) <= 1e-12
