"""n-ary contraction executor: path finding, liveness arena and label bookkeeping on CPU;
device results against the oracle's fold of pairwise binary_einsums on GPU, including CUDA-graph replay."""
import numpy as np
import pytest

from cases import integer_array, random_array
from oracle import contract_path_oracle, rel_frobenius

TOL = {"float32": 1e-5, "complex64": 1e-5, "float64": 1e-12, "complex128": 1e-12}

# (name, [labels per tensor], extents, out)
NETWORKS = [
    ("mps_mpo_transfer", ["awb", "bsc", "wstv", "ate"], dict(a=12, b=12, c=12, e=12, w=4, v=4, s=2, t=2), "evc"),
    ("matrix_chain", ["ij", "jk", "kl", "lm"], dict(i=7, j=9, k=5, l=8, m=6), "im"),
    ("ring", ["ab", "bc", "cd", "da"], dict(a=6, b=5, c=4, d=3), ""),
    ("star_hyper", ["ax", "bx", "cx"], dict(a=5, b=4, c=3, x=6), "abc"),
    ("hyper_kept", ["ax", "bx", "cx"], dict(a=5, b=4, c=3, x=6), "xcab"),
    ("dangling", ["ij", "jk", "kz"], dict(i=4, j=5, k=6, z=7), "i"),
    ("outer", ["ab", "cd", "bx"], dict(a=3, b=4, c=5, d=2, x=6), "dcax"),
    ("peps_patch", ["abcd", "cefg", "bhei", "dgjk", "hl", "fl"], dict(a=3, b=4, c=3, d=2, e=4, f=3, g=2, h=3, i=2, j=3, k=2, l=4), "aijk"),
    ("qubits", ["abcdef", "agbh", "cidj", "ekfl", "gm"], {c: 2 for c in "abcdefghijklm"}, "mhijkl"),
    ("single_permute", ["abc"], dict(a=4, b=5, c=6), "cab"),
    ("single_reduce", ["abc"], dict(a=4, b=5, c=6), "b"),
]


def _make(rng, labels, ext, dt, integer=False):
    gen = integer_array if integer else random_array
    return [gen(rng, tuple(ext[c] for c in ix), dt) for ix in labels]


def _einsum_all(arrays, labels, out):
    m = {}
    for ix in labels:
        for c in ix:
            m.setdefault(c, len(m))
    args = []
    for x, ix in zip(arrays, labels):
        args += [x, [m[c] for c in ix]]
    return np.einsum(*args, [m[c] for c in out])


# ------------------------------------------------------------------------------------------------ CPU
@pytest.mark.parametrize("net", NETWORKS, ids=[n[0] for n in NETWORKS])
def test_program_compiles_and_oracle_fold_matches_full_einsum(net):
    import muscle_b200 as mb
    name, labels, ext, out = net
    I = lambda s: [mb.Index(c) for c in s]
    prog = mb.ContractionProgram([I(ix) for ix in labels], [tuple(ext[c] for c in ix) for ix in labels],
                                 ["complex128"] * len(labels), out=I(out))
    assert len(prog.path) == len(labels) - 1
    assert prog.out == I(out) and prog.result_shape == tuple(ext[c] for c in out)
    # every step's output labels are needed later or in out; the last binary step writes `out` order
    last = [s for s in prog.steps if s["kind"] == "binary"]
    if last:
        assert len(last[-1]["mc"]) == len(out)
    rng = np.random.default_rng(0)
    arrays = _make(rng, labels, ext, "complex128")
    fold = contract_path_oracle(arrays, [list(ix) for ix in labels], list(out), prog.path)
    assert rel_frobenius(fold, _einsum_all(arrays, labels, out)) < 1e-13


def test_arena_liveness_no_overlap():
    import muscle_b200 as mb
    from muscle_b200.network import _Arena
    ar = _Arena()
    a = ar.alloc(1000); b = ar.alloc(5000); c = ar.alloc(300)
    assert a[0] % 256 == 0 and b[0] >= a[0] + a[1] and c[0] >= b[0] + b[1]
    ar.release(*b)
    d = ar.alloc(4000)
    assert d[0] == b[0]                      # reuses the freed block
    ar.release(*a); ar.release(*d); ar.release(*c)
    assert ar.free == [(0, ar.top)]          # everything coalesced
    # a chain of equal-size intermediates needs two buffers, not n
    I = lambda s: [mb.Index(c) for c in s]
    n = 8
    labels = [f"{chr(97 + k)}{chr(98 + k)}" for k in range(n)]
    prog = mb.ContractionProgram([I(ix) for ix in labels], [(64, 64)] * n, ["float64"] * n, out=I("a" + chr(97 + n)),
                                 path=[(0, 1)] + [(n + k, k + 2) for k in range(n - 2)])
    assert prog.arena_bytes <= 2 * 64 * 64 * 8


def test_program_rejects():
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    with pytest.raises(mb.ArgumentError):
        mb.ContractionProgram([I("ij"), I("jk")], [(2, 3), (3, 4)], ["float64"] * 2, out=I("iz"))
    with pytest.raises(mb.DimensionMismatch):
        mb.ContractionProgram([I("ij"), I("jk")], [(2, 3), (4, 4)], ["float64"] * 2)
    with pytest.raises(mb.ArgumentError):
        mb.ContractionProgram([I("ij"), I("jk")], [(2, 3), (3, 4)], ["float64"] * 2, path=[(0, 0)])
    with pytest.raises(mb.ArgumentError):
        mb.ContractionProgram([I("ii")], [(2, 2)], ["float64"])


def test_intermediate_order_puts_next_summed_labels_first():
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    labels = ["awb", "bsc", "wstv", "ate"]
    ext = dict(a=12, b=12, c=12, e=12, w=4, v=4, s=2, t=2)
    prog = mb.ContractionProgram([I(ix) for ix in labels], [tuple(ext[c] for c in ix) for ix in labels],
                                 ["complex128"] * 4, out=I("evc"), path=[(0, 1), (4, 2), (5, 3)])
    o4, o5 = prog.intermediate_orders[4], prog.intermediate_orders[5]
    assert o4[:2] == I("ws")      # summed with W[w,s,t,v] next, in W's memory order
    assert o5[:2] == I("at")      # summed with Ā[a,t,e] next


# ------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize("dt", ["complex128", "complex64", "float64", "float32"])
@pytest.mark.parametrize("net", NETWORKS, ids=[n[0] for n in NETWORKS])
def test_network_parity(net, dt):
    import muscle_b200 as mb
    name, labels, ext, out = net
    I = lambda s: [mb.Index(c) for c in s]
    rng = np.random.default_rng(11)
    for integer in (False, True):
        arrays = _make(rng, labels, ext, dt, integer)
        ts = [mb.Tensor(x, I(ix)).to_device() for x, ix in zip(arrays, labels)]
        got = mb.contract(ts, out=I(out))
        assert got.inds == I(out) and got.on_device
        prog = mb.ContractionProgram([t.inds for t in ts], [t.shape for t in ts], [t.dtype for t in ts], out=I(out))
        wide = [x.astype(np.complex128 if np.dtype(dt).kind == "c" else np.float64) for x in arrays]
        ref = contract_path_oracle(arrays if integer else wide, [list(ix) for ix in labels], list(out), prog.path)
        g = got.to_host().data
        assert g.shape == ref.shape
        if integer:
            assert np.array_equal(g, ref)
        else:
            assert rel_frobenius(g.astype(ref.dtype), ref) <= TOL[dt]


@pytest.mark.gpu
def test_network_mps_mpo_chain_midsize_and_graph_replay():
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    labels = ["awb", "bsc", "wstv", "ate"]
    ext = dict(a=96, b=96, c=96, e=96, w=8, v=8, s=2, t=2)
    rng = np.random.default_rng(2)
    arrays = _make(rng, labels, ext, "complex128")
    ts = [mb.Tensor(x, I(ix)).to_device() for x, ix in zip(arrays, labels)]
    prog = mb.ContractionProgram([t.inds for t in ts], [t.shape for t in ts], [t.dtype for t in ts], out=I("evc"),
                                 path=[(0, 1), (4, 2), (5, 3)])
    ref = contract_path_oracle(arrays, [list(ix) for ix in labels], list("evc"), prog.path)
    got = prog.run(ts)
    assert rel_frobenius(got.to_host().data, ref) <= 1e-12
    h = mb.Handle.get(0)
    h.reset_stats()
    cap = prog.capture(ts)
    assert rel_frobenius(cap.replay().to_host().data, ref) <= 1e-12
    # new contents in the same buffers: the replay sees them
    arrays2 = _make(rng, labels, ext, "complex128")
    for t, x in zip(ts, arrays2):
        t.data.copy_from_host(x)
    ref2 = contract_path_oracle(arrays2, [list(ix) for ix in labels], list("evc"), prog.path)
    out = cap.replay()
    assert rel_frobenius(out.to_host().data, ref2) <= 1e-12
    assert h.stats()["graph_launches"] == 2


@pytest.mark.gpu
def test_network_graph_replay_small_complex64_chain():
    """A launch-bound chain (tiny tensors, tcgen05 never eligible): replay == run."""
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    name, labels, ext, out = NETWORKS[7]
    rng = np.random.default_rng(4)
    arrays = _make(rng, labels, ext, "complex64")
    ts = [mb.Tensor(x, I(ix)).to_device() for x, ix in zip(arrays, labels)]
    prog = mb.ContractionProgram([t.inds for t in ts], [t.shape for t in ts], [t.dtype for t in ts], out=I(out))
    a = prog.run(ts).to_host().data
    cap = prog.capture(ts)
    for _ in range(3):
        b = cap.replay().to_host().data
    assert np.array_equal(a, b)


@pytest.mark.gpu
def test_graph_capture_error_paths():
    """mb200_graph_begin refuses the legacy default stream; a plan that is not cached yet cannot be built while
    capturing (NOT_SUPPORTED -> ArgumentError) and the capture is still closed cleanly."""
    import ctypes as C
    import torch
    import muscle_b200 as mb
    from muscle_b200 import _lib
    h = _lib.Handle.get(0)
    h.set_stream(0)
    with pytest.raises(mb.ArgumentError, match="non-default stream"):
        _lib.check(mb.lib().mb200_graph_begin(h.ptr))
    I = lambda s: [mb.Index(c) for c in s]
    a = mb.Tensor(np.ones((5, 7, 3)), I("xyz")).to_device()
    b = mb.Tensor(np.ones((7, 11, 3)), I("ywz")).to_device()      # a shape no other test has planned
    side = torch.cuda.Stream(device=0)
    with torch.cuda.stream(side):
        h = _lib.Handle.get(0)
        _lib.check(mb.lib().mb200_graph_begin(h.ptr))
        try:
            with pytest.raises(mb.ArgumentError, match="plan miss during graph capture"):
                mb.binary_einsum(mb.BackendB200(), I("xwz"), a, b)
        finally:
            g = C.c_void_p()
            _lib.check(mb.lib().mb200_graph_end(h.ptr, C.byref(g)))
            _lib.check(mb.lib().mb200_graph_destroy(g))
    torch.cuda.synchronize()
    c = mb.binary_einsum(a, b, out=I("xwz"))                         # the handle is usable again
    assert np.array_equal(c.to_host().data, 7.0 * np.ones((5, 11, 3)))


@pytest.mark.gpu
def test_graph_replay_survives_plan_cache_eviction():
    """a captured graph bakes the device pointers of its plans' offset tables into its kernel nodes. The LRU plan
    cache (128 entries per handle) used to cudaFree those tables on eviction, so a later replay gathered and scattered through freed
    memory. Graphs now co-own their plans: after more than 128 OTHER gather-GEMM plans have been built on the same handle (each
    allocating tables, so freed blocks would be reused and overwritten), the replay must still give the right answer."""
    import muscle_b200 as mb
    I = lambda s: [mb.Index(c) for c in s]
    labels = ["awb", "bsc", "wstv", "ate"]
    ext = dict(a=64, b=64, c=64, e=64, w=8, v=8, s=2, t=2)
    rng = np.random.default_rng(5)
    arrays = _make(rng, labels, ext, "complex128")
    ts = [mb.Tensor(x, I(ix)).to_device() for x, ix in zip(arrays, labels)]
    prog = mb.ContractionProgram([t.inds for t in ts], [t.shape for t in ts], [t.dtype for t in ts], out=I("evc"),
                                 path=[(0, 1), (4, 2), (5, 3)])
    ref = contract_path_oracle(arrays, [list(ix) for ix in labels], list("evc"), prog.path)
    cap = prog.capture(ts)
    assert rel_frobenius(cap.replay().to_host().data, ref) <= 1e-12
    h = mb.Handle.get(0)
    built0 = h.stats()["plans_built"]
    # 140 distinct shapes on the tiled path (tables on the device), well past the cache capacity
    for k in range(140):
        m, n, kk = 72 + k, 40 + (k % 7), 48
        a = mb.Tensor(np.ones((m, kk), np.complex128), I("ik")).to_device()
        b = mb.Tensor(np.full((kk, n), 2.0, np.complex128), I("kj")).to_device()
        h.set_path(mb.PATH_GETT_F64)
        c = mb.binary_einsum(a, b, out=I("ij"))
        h.set_path(mb.PATH_AUTO)
    assert np.array_equal(c.to_host().data, np.full((m, n), 2.0 * kk))
    assert h.stats()["plans_built"] - built0 >= 140
    # the captured chain's plans have been evicted from the cache; the graph still owns them
    assert rel_frobenius(cap.replay().to_host().data, ref) <= 1e-12
    arrays2 = _make(rng, labels, ext, "complex128")
    for t, x in zip(ts, arrays2):
        t.data.copy_from_host(x)
    ref2 = contract_path_oracle(arrays2, [list(ix) for ix in labels], list("evc"), prog.path)
    assert rel_frobenius(cap.replay().to_host().data, ref2) <= 1e-12
    # and an un-captured run of the same chain simply rebuilds its plans
    assert rel_frobenius(prog.run(ts).to_host().data, ref2) <= 1e-12
