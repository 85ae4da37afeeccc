"""Multi-GPU `binary_einsum`: one process per GPU, torch.distributed (NCCL over NVLink 5 / NVSwitch on
the box, gloo in CPU tests) for the plumbing.

Mirrors the reference's only distributed path, Dagger block sharding
(ext/MuscleDaggerExt/binary_einsum.jl:64-119):
  * splitting a FREE or BATCH index gives independent output blocks — Dagger's loop over output blocks
    (:88-105). Here: each rank contracts its slab, no collective on the data path.
  * splitting a SUMMED index gives full-size partial outputs that are add-reduced — Dagger's
    `treereduce(AddComputeOp, …)` over the summed blocks (:107-115). Here: `all_reduce(SUM)`.
The per-shard contraction is the ordinary single-GPU `binary_einsum` (Dagger re-enters
`Muscle.binary_einsum` per chunk too, :60-62).
"""
from __future__ import annotations

import numpy as np

from . import _lib
from .backend import BackendB200
from .einsum import binary_einsum, flatten_labels
from .tensor import B200Array, Tensor, _as_index_list


def plan_shard(a: Tensor, b: Tensor, inds_c, nranks: int, rank: int, prefer_sum: bool = False):
    """Which index this rank's work is cut along. Returns (kind, index, begin, end, needs_allreduce);
    kind is one of _lib.SHARD_NONE / SHARD_FREE / SHARD_BATCH / SHARD_SUM."""
    inds_c = _as_index_list(inds_c)
    ma, mb, mc = flatten_labels(a.inds, b.inds, inds_c)
    info = _lib.shard_plan(mc, ma, a.shape, mb, b.shape, nranks, rank, prefer_sum)
    index = None
    if info.kind != _lib.SHARD_NONE:
        for lab, m in list(zip(a.inds, ma)) + list(zip(b.inds, mb)):
            if m == info.mode:
                index = lab
                break
    return info.kind, index, int(info.begin), int(info.end), bool(info.needs_allreduce)


def local_slab(t: Tensor, index, begin: int, end: int) -> Tensor:
    """This rank's slab of a HOST tensor along `index` (the data-distribution step; tensors that do not
    carry the index are replicated). Device tensors are expected to be created per rank already."""
    if index not in t.inds:
        return t
    if t.on_device:
        raise _lib.ArgumentError("local_slab works on host tensors; create device slabs per rank")
    sl = [slice(None)] * t.ndim
    sl[t.dim(index)] = slice(begin, end)
    return Tensor(_lib.fortran(t.data[tuple(sl)]), t.inds)


def _default_contract(inds_c, a, b):
    return binary_einsum(BackendB200(), inds_c, a, b)


def sharded_binary_einsum(a: Tensor, b: Tensor, inds_c, *, group=None, prefer_sum=False, gather=False,
                          contract=_default_contract, fused=True):
    """Contract replicated host (or per-rank device) operands across the ranks of `group`.

    Returns (c_local, info). For a free/batch shard `c_local` is this rank's slab of C (or, with
    gather=True, the full C assembled by all_gather along the split index); for a summed-index slice
    it is the full C after the add-reduction of the partial outputs: the all-reduce FUSED into the contraction
    (`sum_slice_all_reduce`, cross-GPU split-K over NVLink peer memory) whenever the slice runs on the tcgen05
    path and `fused` is left on, else contraction followed by all_reduce(SUM) (NCCL). `contract` exists so the
    CPU (gloo) tests can exercise this host logic without a GPU; the product default is BackendB200.
    """
    import torch
    import torch.distributed as dist

    inds_c = _as_index_list(inds_c)
    nranks = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    kind, index, begin, end, needs_allreduce = plan_shard(a, b, inds_c, nranks, rank, prefer_sum)
    if kind == _lib.SHARD_NONE:
        return contract(inds_c, a, b), (kind, index, begin, end)     # replicas only
    a_loc = local_slab(a, index, begin, end)
    b_loc = local_slab(b, index, begin, end)
    if needs_allreduce and contract is _default_contract and nranks > 1 and dist.get_backend(group) == "nccl":
        # NCCL reduces device buffers: the slices go to this rank's GPU first (host operands are the data-distribution convenience)
        dev = a_loc.data.device if a_loc.on_device else (b_loc.data.device if b_loc.on_device else _lib.current_device())
        a_loc, b_loc = a_loc.to_device(dev), b_loc.to_device(dev)
        if fused:
            try:
                return sum_slice_all_reduce(a_loc, b_loc, inds_c, group=group), (kind, index, begin, end)
            except _lib.ArgumentError:
                pass                              # not on the tcgen05 path (ComplexF64, small shapes): NCCL below
    c_loc = contract(inds_c, a_loc, b_loc)
    if needs_allreduce:
        c_loc = all_reduce_sum(c_loc, group)
    elif gather:
        c_loc = all_gather_along(c_loc, index, group)
    return c_loc, (kind, index, begin, end)


def _as_torch_view(c: Tensor):
    """A torch tensor aliasing C's storage (device: the owning uint8 buffer reinterpreted; host: numpy)."""
    import torch
    if c.on_device:
        owner = c.data._owner
        if owner is None:
            raise _lib.B200Error("device array is not backed by a torch allocation; cannot hand it to NCCL")
        real = torch.float32 if c.dtype in (np.dtype(np.float32), np.dtype(np.complex64)) else torch.float64
        return owner[: c.data.nbytes].view(real)
    arr = c.data
    flat = arr.reshape(-1, order="F") if arr.ndim else arr.reshape(1)
    if not (arr.flags.f_contiguous or arr.ndim <= 1):
        raise _lib.ArgumentError("host C must be column-major contiguous")
    real = np.float32 if arr.dtype in (np.dtype(np.float32), np.dtype(np.complex64)) else np.float64
    return torch.from_numpy(flat.view(real))


def all_reduce_sum(c: Tensor, group=None) -> Tensor:
    """Σ over ranks of the partial outputs, in place (a sum does not care about re/im interleaving, so the
    buffer is reduced as plain reals)."""
    import torch.distributed as dist
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        if c.on_device:
            _lib.Handle.get(c.data.device)       # make sure our launches and NCCL share torch's current stream
        dist.all_reduce(_as_torch_view(c), op=dist.ReduceOp.SUM, group=group)
    return c


def all_gather_along(c: Tensor, index, group=None) -> Tensor:
    """Assemble the full C from per-rank slabs along `index` (host tensors; used by tests and by callers
    that want C replicated — the timed sharded path keeps C sharded)."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return c
    host = c.to_host()
    parts = [None] * dist.get_world_size(group)
    dist.all_gather_object(parts, host.data, group=group)
    full = np.concatenate(parts, axis=host.dim(index))
    return Tensor(_lib.fortran(full), host.inds)


# ---- fused contraction + reduce-scatter over peer memory (summed-index slice) -------------------------------
class _PeerStaging:
    """Two symmetric staging buffers per rank (cudaMalloc'd, exported through CUDA IPC) and the peer mappings of
    every other rank's buffers. Buffers alternate between calls so a fast rank can never overwrite slots an owner
    is still reducing (the next call's barrier orders them)."""

    def __init__(self, device, nbytes, group):
        import ctypes as C
        import torch.distributed as dist
        self.handle = _lib.Handle.get(device)
        L = _lib.lib()
        self.nranks = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.own, exported = [], []
        self.flag_off = (nbytes + 255) // 256 * 256      # nranks ints after the slots: the flag barrier's array
        self.epoch = 0
        for _ in range(2):
            p = C.c_void_p()
            _lib.check(L.mb200_malloc(self.handle.ptr, C.byref(p), self.flag_off + 256))
            _lib.check(L.mb200_memset(self.handle.ptr, C.c_void_p(p.value + self.flag_off), 0, 256))
            self.own.append(int(p.value))
            buf = C.create_string_buffer(64)
            if self.nranks > 1:
                _lib.check(L.mb200_ipc_export(self.handle.ptr, p, buf))
            exported.append(buf.raw)
        gathered = [exported]
        if self.nranks > 1:
            gathered = [None] * self.nranks
            dist.all_gather_object(gathered, exported, group=group)
        self.peers = []
        for b in range(2):
            ptrs = []
            for r in range(self.nranks):
                if r == self.rank:
                    ptrs.append(self.own[b])
                else:
                    q = C.c_void_p()
                    _lib.check(L.mb200_ipc_import(self.handle.ptr, gathered[r][b], C.byref(q)))
                    ptrs.append(int(q.value))
            self.peers.append(ptrs)
        self.turn = 0
        self.handle.synchronize()                        # flag arrays are zero before any peer can signal
        if self.nranks > 1:
            dist.barrier(group=group)


_STAGING: dict = {}


def sum_slice_reduce_scatter(a_loc: Tensor, b_loc: Tensor, inds_c, group=None) -> Tensor:
    """Summed-index slice with the reduction fused into the GEMM epilogue: this rank contracts its K-slice and
    the epilogue stores every output element into the owner rank's staging slot over NVLink peer memory
    (mb200_binary_einsum_scatter); after a stream-ordered cross-rank barrier each rank sums its slots
    (mb200_reduce_slots). Returns this rank's slab of C — the flat column-major range
    [rank*slab, (rank+1)*slab), i.e. reduce-scatter semantics (C stays sharded).
    Raises ArgumentError when the contraction is not on a tensor-core path or C does not split into
    power-of-two slabs; callers then use `all_reduce_sum`."""
    import ctypes as C
    import torch
    import torch.distributed as dist

    inds_c = _as_index_list(inds_c)
    if not (a_loc.on_device and b_loc.on_device):
        raise _lib.ArgumentError("sum_slice_reduce_scatter needs device-resident slices")
    nranks = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    ma, mb, mc = flatten_labels(a_loc.inds, b_loc.inds, inds_c)
    T = np.result_type(a_loc.dtype, b_loc.dtype)
    ext = {}
    for t in (a_loc, b_loc):
        for i, e in zip(t.inds, t.shape):
            ext[i] = e
    shape_c = tuple(ext[i] for i in inds_c)
    numel = int(np.prod(shape_c, dtype=np.int64)) if shape_c else 1
    slab = numel // nranks
    if nranks > _MAX_PEERS or slab * nranks != numel or slab & (slab - 1) or slab == 0:
        raise _lib.ArgumentError("C does not split into power-of-two slabs over the ranks")
    dev = a_loc.data.device
    nbytes = numel * T.itemsize
    key = (dev, nranks, nbytes, id(group))
    st = _STAGING.get(key)
    if st is None:
        st = _STAGING[key] = _PeerStaging(dev, nbytes, group)
    b = st.turn
    st.turn ^= 1
    h = _lib.Handle.get(dev)
    L = _lib.lib()
    arr = (C.c_void_p * nranks)(*st.peers[b])
    _lib.check(L.mb200_binary_einsum_scatter(
        h.ptr, _lib.dtype_enum(T), len(mc), _lib.i32(mc),
        C.c_void_p(a_loc.data.ptr), _lib.dtype_enum(a_loc.dtype), len(ma), _lib.i32(ma), _lib.i64(a_loc.shape), None,
        C.c_void_p(b_loc.data.ptr), _lib.dtype_enum(b_loc.dtype), len(mb), _lib.i32(mb), _lib.i64(b_loc.shape), None,
        arr, nranks, rank, slab.bit_length() - 1))
    # cross-rank barrier without a collective: signal this call's epoch into every rank's flag array (stream-ordered
    # after the GEMM and its peer stores); the slot-sum kernel waits for all nranks entries of the local array
    st.epoch += 1
    flags = (C.c_void_p * nranks)(*[p + st.flag_off for p in st.peers[b]])
    _lib.check(L.mb200_signal_peers(h.ptr, flags, nranks, rank, st.epoch))
    # slab as an array: split C's slowest mode when it divides evenly, else a flat vector
    shaped = bool(shape_c) and shape_c[-1] % nranks == 0
    out = B200Array(shape_c[:-1] + (shape_c[-1] // nranks,) if shaped else (slab,), T, dev)
    _lib.check(L.mb200_reduce_slots_wait(h.ptr, C.c_void_p(out.ptr), C.c_void_p(st.own[b]), _lib.dtype_enum(T), slab, nranks,
                                         C.c_void_p(st.own[b] + st.flag_off), st.epoch))
    if shaped:
        return Tensor(out, inds_c)
    from .tensor import Index
    return Tensor(out, [Index(("flat", tuple(i.tag for i in inds_c)))])


_MAX_PEERS = 8


# ---- fused contraction + ALL-reduce: cross-GPU split-K (mb200_binary_einsum_allreduce) ------------------------------------
class _SymmetricBuffers:
    """One symmetric allocation per rank holding  ws | flags | C  and every rank's mapping of it.

    Plumbing (torch is the allocator / rendezvous, the kernels only see raw pointers):
      * "symm": torch.distributed._symmetric_memory (CUDA VMM handles exchanged over the process group); when the fabric
        supports it the same allocation is also bound to an NVLS multicast object -> `multicast_ptr` (the reducer then uses
        multimem.ld_reduce / multimem.st: the NVSwitch adds and replicates).
      * "ipc": cudaMalloc + CUDA IPC handles (the plumbing of the fused reduce-scatter above); peer loads / stores only.
    MB200_DIST_PLUMBING=symm|ipc forces one; MB200_DIST_MULTICAST=0 ignores an available multicast mapping."""

    def __init__(self, device, ws_bytes, flag_bytes, c_bytes, group):
        import ctypes as C
        import os
        import torch
        import torch.distributed as dist
        self.nranks = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        al = lambda n: (n + 4095) // 4096 * 4096
        self.ws_off, self.flag_off, self.c_off = 0, al(ws_bytes), al(ws_bytes) + al(flag_bytes)
        total = self.c_off + al(c_bytes)
        self.ws_bytes, self.flag_bytes, self.c_bytes = ws_bytes, flag_bytes, c_bytes
        self.epoch = 0
        self.mc = 0
        self.kind = None
        want = os.environ.get("MB200_DIST_PLUMBING", "")
        self.handle = _lib.Handle.get(device)
        if want != "ipc" and self.nranks > 1:
            try:
                import torch.distributed._symmetric_memory as symm_mem
                self._t = symm_mem.empty(total, dtype=torch.uint8, device=f"cuda:{device}")
                self._t.zero_()
                hdl = symm_mem.rendezvous(self._t, group if group is not None else dist.group.WORLD)
                self.bases = [int(p) for p in hdl.buffer_ptrs]
                if os.environ.get("MB200_DIST_MULTICAST", "1") != "0" and int(hdl.multicast_ptr or 0):
                    self.mc = int(hdl.multicast_ptr)
                self._hdl = hdl
                self.kind = "symm"
            except Exception as e:  # noqa: BLE001
                if want == "symm":
                    raise
                self._symm_error = repr(e)[:300]
        if self.kind is None:
            L = _lib.lib()
            p = C.c_void_p()
            _lib.check(L.mb200_malloc(self.handle.ptr, C.byref(p), total))
            _lib.check(L.mb200_memset(self.handle.ptr, p, 0, total))
            own = int(p.value)
            self.bases = [own]
            if self.nranks > 1:
                buf = C.create_string_buffer(64)
                _lib.check(L.mb200_ipc_export(self.handle.ptr, p, buf))
                gathered = [None] * self.nranks
                dist.all_gather_object(gathered, buf.raw, group=group)
                self.bases = []
                for r in range(self.nranks):
                    if r == self.rank:
                        self.bases.append(own)
                    else:
                        q = C.c_void_p()
                        _lib.check(L.mb200_ipc_import(self.handle.ptr, gathered[r], C.byref(q)))
                        self.bases.append(int(q.value))
            self.kind = "ipc"
        self.handle.synchronize()                        # flags are zero before any peer can signal
        try:
            import torch
            torch.cuda.synchronize(device)
        except Exception:  # noqa: BLE001
            pass
        if self.nranks > 1:
            dist.barrier(group=group)

    def c_owner(self, device):
        """uint8 torch tensor aliasing this rank's C region (the `_owner` a B200Array expects)."""
        import torch
        if self.kind == "symm":
            return self._t[self.c_off: self.c_off + max(self.c_bytes, 1)]
        if getattr(self, "_c_view", None) is None:
            class _Raw:
                pass
            raw = _Raw()
            raw.__cuda_array_interface__ = {"shape": (max(self.c_bytes, 1),), "typestr": "|u1", "version": 2,
                                            "data": (self.bases[self.rank] + self.c_off, False)}
            self._c_view = torch.as_tensor(raw, device=f"cuda:{device}")
            self._c_view._mb200_keepalive = self
        return self._c_view

    def comm(self, epoch) -> "_lib.Comm":
        cm = _lib.Comm()
        cm.nranks, cm.rank, cm.epoch = self.nranks, self.rank, epoch
        for r in range(self.nranks):
            cm.ws[r] = self.bases[r] + self.ws_off
            cm.flags[r] = self.bases[r] + self.flag_off
            cm.c[r] = self.bases[r] + self.c_off
        cm.mc_ws = (self.mc + self.ws_off) if self.mc else None
        cm.mc_c = (self.mc + self.c_off) if self.mc else None
        cm.ws_bytes, cm.flag_bytes = self.ws_bytes, self.flag_bytes
        return cm


_ALLREDUCE: dict = {}
_ALLREDUCE_CAP = 8


def allreduce_workspace(a_loc: Tensor, b_loc: Tensor, inds_c, nranks: int):
    """(ws_bytes, flag_bytes) of the fused all-reduce for this contraction; ArgumentError when it is not on the tcgen05 path."""
    import ctypes as C
    inds_c = _as_index_list(inds_c)
    ma, mb, mc = flatten_labels(a_loc.inds, b_loc.inds, inds_c)
    T = np.result_type(a_loc.dtype, b_loc.dtype)
    ws, fl = C.c_size_t(), C.c_size_t()
    h = _lib.Handle.get(a_loc.data.device if a_loc.on_device else None)
    _lib.check(_lib.lib().mb200_allreduce_workspace(
        h.ptr, _lib.dtype_enum(T), len(mc), _lib.i32(mc),
        _lib.dtype_enum(a_loc.dtype), len(ma), _lib.i32(ma), _lib.i64(a_loc.shape),
        _lib.dtype_enum(b_loc.dtype), len(mb), _lib.i32(mb), _lib.i64(b_loc.shape),
        nranks, C.byref(ws), C.byref(fl)))
    return int(ws.value), int(fl.value)


def sum_slice_all_reduce(a_loc: Tensor, b_loc: Tensor, inds_c, group=None, phases=7, bump_epoch=True) -> Tensor:
    """Summed-index slice with the all-reduce fused into the contraction (all-reduce semantics: every rank ends with the
    full C, bit-identical on all ranks) - Dagger's `treereduce(AddComputeOp, ...)` over the summed blocks
    (ext/MuscleDaggerExt/binary_einsum.jl:107-115) without a separate collective pass: see mb200_binary_einsum_allreduce.

    The returned Tensor aliases a symmetric buffer that the NEXT call with the same signature overwrites (and that is
    released once `_ALLREDUCE_CAP` other signatures have been used); copy it (`permutedims`, `to_host`) if it must outlive that. Raises ArgumentError when the contraction is not on the
    tcgen05 path; callers then use `all_reduce_sum` on the partial outputs."""
    import ctypes as C
    import torch.distributed as dist

    inds_c = _as_index_list(inds_c)
    if not (a_loc.on_device and b_loc.on_device):
        raise _lib.ArgumentError("sum_slice_all_reduce needs device-resident slices")
    nranks = dist.get_world_size(group) if dist.is_initialized() else 1
    if nranks > _MAX_PEERS:
        raise _lib.ArgumentError(f"at most {_MAX_PEERS} ranks")
    ma, mb, mc = flatten_labels(a_loc.inds, b_loc.inds, inds_c)
    T = np.result_type(a_loc.dtype, b_loc.dtype)
    ext = {}
    for t in (a_loc, b_loc):
        for i, e in zip(t.inds, t.shape):
            ext[i] = e
    shape_c = tuple(ext[i] for i in inds_c)
    numel = int(np.prod(shape_c, dtype=np.int64)) if shape_c else 1
    dev = a_loc.data.device
    key = (dev, nranks, id(group), T.str, tuple(ma), tuple(a_loc.shape), tuple(mb), tuple(b_loc.shape), tuple(mc))
    st = _ALLREDUCE.get(key)
    if st is None:
        ws_bytes, flag_bytes = allreduce_workspace(a_loc, b_loc, inds_c, nranks)
        # at most _ALLREDUCE_CAP signatures keep their symmetric buffers (workspace + C each); the oldest goes first - every rank
        # makes the same calls in the same order, so every rank drops the same entry at the same point
        while len(_ALLREDUCE) >= _ALLREDUCE_CAP:
            old = next(iter(_ALLREDUCE))
            _lib.Handle.get(dev).synchronize()
            del _ALLREDUCE[old]
        st = _ALLREDUCE[key] = _SymmetricBuffers(dev, ws_bytes, flag_bytes, numel * T.itemsize, group)
    if bump_epoch:                                   # False: a later phase of the call that raised the epoch (diagnostics, tests)
        st.epoch += 1
    h = _lib.Handle.get(dev)
    cm = st.comm(st.epoch)
    _lib.check(_lib.lib().mb200_binary_einsum_allreduce(
        h.ptr, _lib.dtype_enum(T), len(mc), _lib.i32(mc),
        C.c_void_p(a_loc.data.ptr), _lib.dtype_enum(a_loc.dtype), len(ma), _lib.i32(ma), _lib.i64(a_loc.shape), None,
        C.c_void_p(b_loc.data.ptr), _lib.dtype_enum(b_loc.dtype), len(mb), _lib.i32(mb), _lib.i64(b_loc.shape), None,
        C.byref(cm), phases))
    out = B200Array(shape_c, T, dev, _owner=st.c_owner(dev), _ptr=st.bases[st.rank] + st.c_off)
    return Tensor(out, inds_c)


def allreduce_plumbing_info() -> list:
    """What the cached fused-all-reduce buffers run on: [(kind, multicast?)] - for bench / test reports."""
    return [(s.kind, bool(s.mc)) for s in _ALLREDUCE.values()]
