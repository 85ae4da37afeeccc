"""Soak test (torchrun, N >= 2) of the fused contraction + all-reduce: many back-to-back calls on CHANGING data, every result compared
on the device with contraction-then-NCCL of the same operands (rel. Frobenius <= 1e-5) and its checksum compared across ranks
(bit-identical). A stale read of a partial unit (a visibility race between the epilogue's stores, the flag and the reducer's
multimem.ld_reduce) would show up as a mismatch against the NCCL result of the current data."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from muscle_b200 import B200Array, Index, Tensor, binary_einsum  # noqa: E402
from muscle_b200 import dist as mdist  # noqa: E402

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 300
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
I = lambda s: [Index(c) for c in s]
n, hl = 8, max(1, 8 // world)
g = torch.Generator(device=f"cuda:{local}"); g.manual_seed(77 + rank)
fa = torch.rand(2 * n ** 7 * hl, dtype=torch.float32, device=f"cuda:{local}", generator=g) * 2 - 1
fb = torch.rand(2 * n ** 7 * hl, dtype=torch.float32, device=f"cuda:{local}", generator=g) * 2 - 1
A = Tensor(B200Array.from_torch(fa, [n] * 7 + [hl], "complex64"), I("aebfcgdh"))
B = Tensor(B200Array.from_torch(fb, [hl] + [n] * 7, "complex64"), I("hpgqfres"))
ic = I("srqpdcba")
numel = n ** 8
bad = 0
worst = 0.0
for it in range(iters):
    fa.mul_(-1.0 if it % 2 else 0.5 + (it % 7) * 0.25)        # new data every call (sign flips and rescales: stale partials cannot match)
    ref = binary_einsum(A, B, out=ic)
    mdist.all_reduce_sum(ref)
    got = mdist.sum_slice_all_reduce(A, B, ic)
    r = ref.data._owner[: 8 * numel].view(torch.float32)
    x = got.data._owner[: 8 * numel].view(torch.float32)
    err = float(torch.linalg.vector_norm(x - r) / torch.linalg.vector_norm(r))
    csum = torch.stack([x.double().sum(), x.double().abs().sum(), (x.double() * (torch.arange(1, 2 * numel + 1, device=x.device, dtype=torch.float64) % 97)).sum()])
    allc = [torch.empty_like(csum) for _ in range(world)]
    dist.all_gather(allc, csum)
    same = all(bool(torch.equal(allc[0], c)) for c in allc)
    worst = max(worst, err)
    if not (err <= 1e-5) or not same:          # a non-finite error counts as a mismatch
        bad += 1
        if rank == 0 and bad <= 5:
            print(f"SOAK mismatch at iteration {it}: err {err:.3e} identical_across_ranks {same}", flush=True)
tot = torch.tensor([bad], device=f"cuda:{local}")
dist.all_reduce(tot)
if rank == 0:
    print(f"SOAK world={world} iterations={iters} mismatches={int(tot.item())} worst_rel_err={worst:.3e} plumbing={mdist.allreduce_plumbing_info()}", flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(1 if int(tot.item()) else 0)
