"""n-ary contraction executor: a chain of `binary_einsum`s along a contraction path, the way
Muscle's callers use the hot path (`binary_einsum(binary_einsum(Θ, U), V)`-style chains in
src/Operations/simple_update.jl:51-80 and test/integration/reactant.jl:111; paths come from EinExprs,
src/Muscle.jl:3).

What the executor adds over calling `binary_einsum` in a loop:
  * every intermediate stays in HBM, in ONE arena sized by liveness analysis (an intermediate's bytes are reused
    as soon as its consumer has run) — no per-step allocation, no host round trips;
  * the label order of each intermediate is chosen for its CONSUMER: labels that the next step sums go first, in the
    partner operand's memory order, so both operands of the next step walk their summed modes as one contiguous
    run (K-major on both sides);
  * labels that are dangling at a step (in one operand, needed by nobody else) are pre-reduced with
    `mb200_unary_einsum`, labels shared by more than two tensors stay as batch (hyper) labels until their last use;
  * all launches are stream-ordered with no synchronisation, and a program over fixed buffers can be captured into
    a CUDA graph (`mb200_graph_*`) and replayed with one launch — the fix for launch-bound networks of small tensors.
Everything runs through the C ABI; there is no CPU path here.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _lib
from ._lib import ArgumentError
from .tensor import B200Array, Index, Tensor, _as_index_list

_ALIGN = 256


def _numel(shape):
    return int(np.prod(shape, dtype=np.int64)) if len(shape) else 1


def find_path(inds_list, sizes, out):
    """Greedy pairwise path (SSA ids): repeatedly contract the pair that minimises size(result) − size(a) − size(b)
    among pairs sharing a label; disconnected components are joined by outer products at the end, smallest first."""
    live = {i: list(ix) for i, ix in enumerate(inds_list)}
    nxt = len(inds_list)
    path = []

    def size_of(ix):
        return float(np.prod([sizes[l] for l in ix], dtype=np.float64)) if ix else 1.0

    def result_inds(a, b):
        others = set(out)
        for k, ix in live.items():
            if k != a and k != b:
                others.update(ix)
        seen, res = set(), []
        for l in live[a] + live[b]:
            if l in others and l not in seen:
                seen.add(l)
                res.append(l)
        return res

    while len(live) > 1:
        best = None
        keys = sorted(live)
        for ai, a in enumerate(keys):
            sa = set(live[a])
            for b in keys[ai + 1:]:
                if not sa.intersection(live[b]):
                    continue
                r = result_inds(a, b)
                cost = size_of(r) - size_of(live[a]) - size_of(live[b])
                if best is None or cost < best[0]:
                    best = (cost, a, b, r)
        if best is None:   # only outer products are left
            a, b = sorted(keys, key=lambda k: size_of(live[k]))[:2]
            best = (0.0, a, b, result_inds(a, b))
        _, a, b, r = best
        path.append((a, b))
        del live[a], live[b]
        live[nxt] = r
        nxt += 1
    return path


class _Arena:
    """First-fit offsets with coalescing free blocks; grows at the end. Offsets are bytes, 256-aligned."""

    def __init__(self):
        self.free = []      # sorted (offset, size)
        self.top = 0

    def alloc(self, nbytes):
        nbytes = max(_ALIGN, (nbytes + _ALIGN - 1) // _ALIGN * _ALIGN)
        for k, (off, sz) in enumerate(self.free):
            if sz >= nbytes:
                if sz == nbytes:
                    self.free.pop(k)
                else:
                    self.free[k] = (off + nbytes, sz - nbytes)
                return off, nbytes
        if self.free and self.free[-1][0] + self.free[-1][1] == self.top:   # extend the trailing free block
            off, sz = self.free.pop()
            self.top = off + nbytes
            return off, nbytes
        off = self.top
        self.top += nbytes
        return off, nbytes

    def release(self, off, nbytes):
        self.free.append((off, nbytes))
        self.free.sort()
        merged = []
        for o, s in self.free:
            if merged and merged[-1][0] + merged[-1][1] == o:
                merged[-1] = (merged[-1][0], merged[-1][1] + s)
            else:
                merged.append((o, s))
        self.free = merged


class ContractionProgram:
    """A compiled n-ary contraction: label bookkeeping, path, intermediate label orders and arena offsets are fixed
    at construction; `run(tensors)` only enqueues kernels.

    inds_list : one label list per input tensor (Index or anything hashable)
    shapes    : one shape per input tensor
    dtypes    : one numpy dtype per input tensor
    out       : labels of the result (default: labels that appear exactly once, in order of appearance)
    path      : list of (i, j) SSA pairs (inputs are 0..n-1, the k-th step produces n+k); default `find_path`
    """

    def __init__(self, inds_list, shapes, dtypes, out=None, path=None):
        self.n = len(inds_list)
        if self.n < 1:
            raise ArgumentError("need at least one tensor")
        self.inds = [_as_index_list(ix) for ix in inds_list]
        self.shapes = [tuple(int(s) for s in sh) for sh in shapes]
        self.dtypes = [np.dtype(d) for d in dtypes]
        self.sizes = {}
        count = {}
        for ix, sh in zip(self.inds, self.shapes):
            if len(ix) != len(sh):
                raise ArgumentError("one label per dimension")
            if len(set(ix)) != len(ix):
                raise ArgumentError("a label repeated inside one tensor: reduce it with unary_einsum first")
            for l, e in zip(ix, sh):
                if self.sizes.setdefault(l, e) != e:
                    raise _lib.DimensionMismatch(f"label {l!r} has extents {self.sizes[l]} and {e}")
                count[l] = count.get(l, 0) + 1
        if out is None:
            out = [l for ix in self.inds for l in ix if count[l] == 1]
        self.out = _as_index_list(out)
        for l in self.out:
            if l not in self.sizes:
                raise ArgumentError(f"output label {l!r} is in no tensor")
        if len(set(self.out)) != len(self.out):
            raise ArgumentError("repeated output labels are not supported")
        self.path = [tuple(p) for p in (path if path is not None else find_path(self.inds, self.sizes, self.out))]
        if len(self.path) != self.n - 1:
            raise ArgumentError(f"a path over {self.n} tensors has {self.n - 1} steps, got {len(self.path)}")
        self._compile()

    # ------------------------------------------------------------------------------------------- compile
    def _compile(self):
        n = self.n
        modeid = {l: k for k, l in enumerate(self.sizes)}
        slot_inds = {i: list(self.inds[i]) for i in range(n)}
        slot_dtype = {i: self.dtypes[i] for i in range(n)}
        consumer = {}
        for t, (a, b) in enumerate(self.path):
            for s in (a, b):
                if s in consumer or s >= n + t or s < 0 or a == b:
                    raise ArgumentError(f"bad path step {t}: ({a}, {b})")
                consumer[s] = t
        live = set(range(n))
        self.steps = []     # dicts: kind, ins, out, modes..., shapes
        self.flops = 0.0
        arena = _Arena()
        slot_mem = {}       # slot -> (offset, nbytes) for intermediates
        nslot = [n + len(self.path)]   # extra slots for pre-reduced operands

        def needed_elsewhere(label, excl):
            if label in self.out:
                return True
            return any(label in slot_inds[s] for s in live if s not in excl)

        def alloc(slot, shape, dtype):
            slot_mem[slot] = arena.alloc(_numel(shape) * np.dtype(dtype).itemsize)

        def free(slot):
            if slot in slot_mem and slot_mem[slot] is not None:
                arena.release(*slot_mem[slot])

        for t, (a, b) in enumerate(self.path):
            res = n + t
            ops = []
            for s, other in ((a, b), (b, a)):
                # dangling labels: only in this operand and needed by nobody else -> sum them away first
                dang = [l for l in slot_inds[s] if l not in slot_inds[other] and not needed_elsewhere(l, (a, b))]
                if dang:
                    keep = [l for l in slot_inds[s] if l not in dang]
                    new = nslot[0]
                    nslot[0] += 1
                    shape = tuple(self.sizes[l] for l in keep)
                    alloc(new, shape, slot_dtype[s])
                    self.steps.append(dict(kind="unary", x=s, y=new, mx=[modeid[l] for l in slot_inds[s]],
                                           my=[modeid[l] for l in keep], xshape=tuple(self.sizes[l] for l in slot_inds[s]),
                                           yshape=shape, dtype=slot_dtype[s]))
                    slot_inds[new] = keep
                    slot_dtype[new] = slot_dtype[s]
                    consumer[new] = t
                    live.add(new)
                    live.discard(s)
                    free(s)
                    ops.append(new)
                else:
                    ops.append(s)
            a2, b2 = ops
            ia, ib = slot_inds[a2], slot_inds[b2]
            kept = [l for l in dict.fromkeys(ia + ib) if needed_elsewhere(l, (a2, b2))]
            last = t == len(self.path) - 1
            if last:
                if set(kept) != set(self.out):
                    raise ArgumentError("the path does not produce the requested output labels")
                order = list(self.out)
            else:
                order = self._order_for_consumer(kept, res, consumer, slot_inds)
            T = np.result_type(slot_dtype[a2], slot_dtype[b2])
            shape = tuple(self.sizes[l] for l in order)
            if not last:
                alloc(res, shape, T)
            self.steps.append(dict(kind="binary", a=a2, b=b2, c=res, ma=[modeid[l] for l in ia], mb=[modeid[l] for l in ib],
                                   mc=[modeid[l] for l in order], ashape=tuple(self.sizes[l] for l in ia),
                                   bshape=tuple(self.sizes[l] for l in ib), cshape=shape, dtype=T,
                                   da=slot_dtype[a2], db=slot_dtype[b2]))
            labels = set(ia) | set(ib)
            self.flops += (8.0 if T.kind == "c" else 2.0) * float(np.prod([self.sizes[l] for l in labels], dtype=np.float64))
            slot_inds[res] = order
            slot_dtype[res] = T
            live.discard(a2)
            live.discard(b2)
            live.add(res)
            free(a2)
            free(b2)
        if self.n == 1:     # a single tensor: permute / reduce to `out`
            ix = slot_inds[0]
            self.steps.append(dict(kind="unary", x=0, y=1, mx=[modeid[l] for l in ix], my=[modeid[l] for l in self.out],
                                   xshape=self.shapes[0], yshape=tuple(self.sizes[l] for l in self.out), dtype=self.dtypes[0]))
            slot_dtype[1] = self.dtypes[0]
            self.result_slot = 1
        else:
            self.result_slot = n + len(self.path) - 1
        self.result_dtype = slot_dtype[self.result_slot]
        self.result_shape = tuple(self.sizes[l] for l in self.out)
        self.slot_mem = slot_mem
        self.arena_bytes = arena.top
        self.intermediate_orders = {s: slot_inds[s] for s in slot_mem}
        self._arena = None
        self._graph = None

    def _order_for_consumer(self, kept, res, consumer, slot_inds):
        """Label order of an intermediate: the labels its consumer step will sum go first (K-major operand), in the
        partner's memory order when the partner is already laid out; the rest keeps a's-then-b's order (the
        reference's default, binary_einsum.jl:38-41). A wrong guess only costs layout, never correctness."""
        t2 = consumer.get(res)
        if t2 is None:
            return kept
        pa, pb = self.path[t2]
        partner = pb if pa == res else pa
        pinds = slot_inds.get(partner)
        if pinds is None:      # the partner is an intermediate that does not exist yet
            return kept

        def alive_after(s):
            c = consumer.get(s)
            return c is None or c > t2

        others = [ix for s, ix in slot_inds.items() if s not in (partner, res) and alive_after(s)]
        soon = [l for l in pinds if l in kept and l not in self.out and not any(l in ix for ix in others)]
        return soon + [l for l in kept if l not in soon]

    # --------------------------------------------------------------------------------------------- run
    def _ensure_arena(self, device):
        if self._arena is None or self._arena.device != device:
            self._arena = _ByteBuffer(max(self.arena_bytes, 1), device)
        return self._arena

    def _check_inputs(self, tensors):
        if len(tensors) != self.n:
            raise ArgumentError(f"expected {self.n} tensors, got {len(tensors)}")
        dev = None
        for k, t in enumerate(tensors):
            if not isinstance(t, Tensor) or not t.on_device:
                raise ArgumentError("ContractionProgram.run takes device-resident Tensors (Tensor.to_device())")
            if t.inds != self.inds[k] or t.shape != self.shapes[k] or t.dtype != self.dtypes[k]:
                raise ArgumentError(f"tensor {k} does not match the program ({t.inds}, {t.shape}, {t.dtype})")
            if dev is None:
                dev = t.data.device
            elif t.data.device != dev:
                raise ArgumentError("all tensors must live on one device")
        return dev

    def _enqueue(self, tensors, out_arr, dev):
        L = _lib.lib()
        h = _lib.Handle.get(dev)
        base = self._ensure_arena(dev).ptr

        def ptr(slot):
            if slot < self.n:
                return tensors[slot].data.ptr
            if slot == self.result_slot:
                return out_arr.ptr
            return base + self.slot_mem[slot][0]

        for st in self.steps:
            if st["kind"] == "unary":
                e = _lib.dtype_enum(st["dtype"])
                _lib.check(L.mb200_unary_einsum(h.ptr, C.c_void_p(ptr(st["y"])), e, len(st["my"]), _lib.i32(st["my"]), None,
                                                C.c_void_p(ptr(st["x"])), e, len(st["mx"]), _lib.i32(st["mx"]),
                                                _lib.i64(st["xshape"]), None))
            else:
                _lib.check(L.mb200_binary_einsum(
                    h.ptr, C.c_void_p(ptr(st["c"])), _lib.dtype_enum(st["dtype"]), len(st["mc"]), _lib.i32(st["mc"]), None,
                    C.c_void_p(ptr(st["a"])), _lib.dtype_enum(st["da"]), len(st["ma"]), _lib.i32(st["ma"]), _lib.i64(st["ashape"]), None,
                    C.c_void_p(ptr(st["b"])), _lib.dtype_enum(st["db"]), len(st["mb"]), _lib.i32(st["mb"]), _lib.i64(st["bshape"]), None))

    def run(self, tensors, out: Tensor | None = None) -> Tensor:
        """Enqueue the whole chain on the current stream (no synchronisation) and return the result tensor."""
        dev = self._check_inputs(tensors)
        if out is None:
            out = Tensor(B200Array(self.result_shape, self.result_dtype, dev), self.out)
        elif not out.on_device or out.inds != self.out or out.shape != self.result_shape or out.dtype != self.result_dtype:
            raise ArgumentError("`out` does not match the program's result")
        self._enqueue(tensors, out.data, dev)
        return out

    def capture(self, tensors, out: Tensor | None = None) -> "CapturedProgram":
        """Run once (builds and caches every plan), then capture the same sequence on the same buffers into a CUDA
        graph. `replay()` re-executes it with one launch; the inputs' CONTENTS may change between replays, their
        addresses may not."""
        import torch
        dev = self._check_inputs(tensors)
        out = self.run(tensors, out)
        torch.cuda.synchronize(dev)
        side = torch.cuda.Stream(device=dev)
        with torch.cuda.stream(side):
            h = _lib.Handle.get(dev)       # follows torch's current (side) stream
            _lib.check(_lib.lib().mb200_graph_begin(h.ptr))
            try:
                self._enqueue(tensors, out.data, dev)
            finally:
                g = C.c_void_p()
                st = _lib.lib().mb200_graph_end(h.ptr, C.byref(g))
            _lib.check(st)
        return CapturedProgram(self, g, out, list(tensors), dev)


class _ByteBuffer:
    """Raw device bytes for the arena (torch's allocator when present, like B200Array)."""

    def __init__(self, nbytes, device):
        self._raw = B200Array(((int(nbytes) + 3) // 4,), np.float32, device)
        self.ptr = self._raw.ptr
        self.device = self._raw.device
        self.nbytes = int(nbytes)


class CapturedProgram:
    def __init__(self, program, graph, out, tensors, device):
        self.program, self._g, self.out, self._keep, self.device = program, graph, out, tensors, device

    def replay(self) -> Tensor:
        h = _lib.Handle.get(self.device)
        _lib.check(_lib.lib().mb200_graph_launch(h.ptr, self._g))
        return self.out

    def __del__(self):
        try:
            if self._g:
                _lib.lib().mb200_graph_destroy(self._g)
        except Exception:
            pass


def contract(tensors, out=None, path=None) -> Tensor:
    """contract([t1, t2, ...]; out, path): n-ary contraction of device-resident tensors along `path` (default: a
    greedy path). Equivalent to folding `binary_einsum` along the path with every intermediate kept on the device."""
    prog = ContractionProgram([t.inds for t in tensors], [t.shape for t in tensors], [t.dtype for t in tensors],
                              out=out, path=path)
    return prog.run(list(tensors))
