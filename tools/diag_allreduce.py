"""Diagnostic (torchrun, N >= 2): where the time of the fused contraction + all-reduce goes. Config 5 slice per rank.
  plain      : binary_einsum only (packs + GEMM, permuting epilogue)
  contract   : phases = CONTRACT (dist-mode GEMM: partial units + flags, no reducer)
  serial     : phases = CONTRACT, then phases = REDUCE | WAIT (reducer after the GEMM, same stream: no overlap by construction)
  fused      : phases = 7 (reducer on the side stream, concurrent)
  nccl       : binary_einsum + all_reduce"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import muscle_b200 as mb  # noqa: E402
from muscle_b200 import B200Array, Index, Tensor, binary_einsum  # noqa: E402
from muscle_b200 import dist as mdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
I = lambda s: [Index(c) for c in s]
n = 8
hl = max(1, n // world)


def rnd(shape, seed):
    g = torch.Generator(device=f"cuda:{local}"); g.manual_seed(seed)
    t = torch.rand(2 * int(np.prod(shape)), dtype=torch.float32, device=f"cuda:{local}", generator=g) * 2 - 1
    return B200Array.from_torch(t, shape, "complex64")


A = Tensor(rnd([n] * 7 + [hl], 5000 + rank), I("aebfcgdh"))
B = Tensor(rnd([hl] + [n] * 7, 5100 + rank), I("hpgqfres"))
ic = I("srqpdcba")


def timed(fn, iters=30):
    for _ in range(3):
        fn()
    torch.cuda.synchronize(); dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / iters], device=f"cuda:{local}", dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.barrier()
    return float(t.item())


def serial():
    mdist.sum_slice_all_reduce(A, B, ic, phases=1)
    mdist.sum_slice_all_reduce(A, B, ic, phases=6, bump_epoch=False)


res = {}
res["plain"] = timed(lambda: binary_einsum(A, B, out=ic))
res["nccl"] = timed(lambda: mdist.all_reduce_sum(binary_einsum(A, B, out=ic)))
res["fused"] = timed(lambda: mdist.sum_slice_all_reduce(A, B, ic))
res["serial"] = timed(serial)
# contract-only: the flags of the next call's epoch are raised but nobody consumes them - harmless, epochs only grow
res["contract"] = timed(lambda: mdist.sum_slice_all_reduce(A, B, ic, phases=1))
tl = ""
if os.environ.get("MB200_DIST_TIMELINE"):
    import ctypes as C
    from muscle_b200 import _lib
    for rep in range(3):
        torch.cuda.synchronize(); dist.barrier()
        mdist.sum_slice_all_reduce(A, B, ic)
        out = (C.c_ulonglong * 8)()
        _lib.check(_lib.lib().mb200_dist_timeline(_lib.Handle.get(local).ptr, out))
        t0 = out[0]
        tl += " | us since GEMM start: gemm_end %.1f reducer_start %.1f first_unit_ready %.1f reducer_end %.1f all_done %.1f" % tuple(
            (int(out[i]) - int(t0)) / 1e3 for i in (1, 2, 3, 4, 5))
if rank == 0:
    print("DIAG world=%d reducer_sms=%s plumbing=%s : " % (world, os.environ.get("MB200_DIST_REDUCER_SMS", "default") + ("" if os.environ.get("MB200_DIST_OVERLAP", "1") != "0" else " NO-OVERLAP"), mdist.allreduce_plumbing_info())
          + "  ".join(f"{k} {v:.4f} ms" for k, v in res.items()) + tl, flush=True)
dist.barrier()
dist.destroy_process_group()
