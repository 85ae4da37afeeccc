"""Worker of tests/test_multi_gpu.py: one process per GPU (torchrun), NCCL. Every multi-GPU data path is checked against the
oracle ON HARDWARE: free-index shard, batch shard, summed-index slice + NCCL all-reduce, fused peer-memory reduce-scatter,
fused all-reduce (cross-GPU split-K). Mirrors test/integration/dagger.jl:12-31 (`collect(block_c) ≈ c`). Prints one JSON line
per rank-0 check; exit code != 0 on any failure."""
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    import torch
    import torch.distributed as dist
    import muscle_b200 as mb
    from muscle_b200 import Index, Tensor, _lib, binary_einsum
    from muscle_b200 import dist as mdist
    from cases import random_array
    from oracle import binary_einsum_base, binary_einsum_general, rel_frobenius

    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    I = lambda s: [Index(c) for c in s]
    report, ok = {}, True

    def check(name, err, tol):
        nonlocal ok
        errs = [None] * world
        dist.all_gather_object(errs, float(err))
        good = all(e <= tol for e in errs)
        ok = ok and good
        report[name] = {"rel_frobenius_per_rank": errs, "tolerance": tol, "ok": good}

    # --- free-index shard (config-4 pattern, ComplexF64): replicated host operands, gather=True assembles the full C
    ext = dict(a=8, b=6, c=4, d=8, e=6, f=4, g=8, h=6, i=2 * world)
    ia, ib, ic = "adbecf", "fgdhei", "abcghi"
    A = random_array(np.random.default_rng(1), tuple(ext[x] for x in ia), "complex128")
    B = random_array(np.random.default_rng(2), tuple(ext[x] for x in ib), "complex128")
    ref = binary_einsum_base(list(ic), A, list(ia), B, list(ib))
    c_full, info = mdist.sharded_binary_einsum(Tensor(A, I(ia)), Tensor(B, I(ib)), I(ic), gather=True,
                                               contract=lambda inds, a, b: binary_einsum(mb.BackendB200(), inds, a.to_device(local), b.to_device(local)))
    assert info[0] == _lib.SHARD_FREE and info[1] == Index("i"), info
    check("free_index_shard_c128", rel_frobenius(c_full.to_host().data, ref), 1e-12)

    # --- batch shard (config-3 pattern, ComplexF32)
    ext = dict(l=64, k=8, b=4, m=64, q=4, r=32, z=2 * world)
    ia, ib, ic = "lkbmz", "mkqrz", "lbqrz"
    A = random_array(np.random.default_rng(3), tuple(ext[x] for x in ia), "complex64")
    B = random_array(np.random.default_rng(4), tuple(ext[x] for x in ib), "complex64")
    ref = binary_einsum_general(list(ic), A.astype(np.complex128), list(ia), B.astype(np.complex128), list(ib))
    c_full, info = mdist.sharded_binary_einsum(Tensor(A, I(ia)), Tensor(B, I(ib)), I(ic), gather=True,
                                               contract=lambda inds, a, b: binary_einsum(mb.BackendB200(), inds, a.to_device(local), b.to_device(local)))
    assert info[0] == _lib.SHARD_BATCH and info[1] == Index("z"), info
    check("batch_shard_c64", rel_frobenius(c_full.to_host().data.astype(np.complex128), ref), 1e-5)

    # --- summed-index slice (config-5 pattern, ComplexF32 on the tcgen05 path): NCCL all-reduce, fused reduce-scatter, fused all-reduce
    n = 8
    ia, ib, ic = "aebfcgdh", "hpgqfres", "srqpdcba"
    ext5 = {x: n for x in "abcdefgpqrs"}
    ext5["h"] = world * (n // world if world <= n else 1)
    A = random_array(np.random.default_rng(5), tuple(ext5[x] for x in ia), "complex64")
    B = random_array(np.random.default_rng(6), tuple(ext5[x] for x in ib), "complex64")
    ref = binary_einsum_base(list(ic), A.astype(np.complex128), list(ia), B.astype(np.complex128), list(ib))
    dev_contract = lambda inds, a, b: binary_einsum(mb.BackendB200(), inds, a.to_device(local), b.to_device(local))
    c_ar, info = mdist.sharded_binary_einsum(Tensor(A, I(ia)), Tensor(B, I(ib)), I(ic), prefer_sum=True, contract=dev_contract)
    assert info[0] == _lib.SHARD_SUM and info[1] == Index("h"), info
    check("sum_slice_nccl_all_reduce_c64", rel_frobenius(c_ar.to_host().data.astype(np.complex128), ref), 1e-5)

    # the product default of a summed-index slice: the all-reduce fused into the contraction (host operands in, full C out)
    h5 = mb.Handle.get(local)
    h5.reset_stats()
    c_def, info = mdist.sharded_binary_einsum(Tensor(A, I(ia)), Tensor(B, I(ib)), I(ic), prefer_sum=True)
    assert info[0] == _lib.SHARD_SUM and c_def.on_device and len(mdist._ALLREDUCE) == 1, (info, mdist._ALLREDUCE)
    check("sum_slice_default_is_fused_all_reduce_c64", rel_frobenius(c_def.to_host().data.astype(np.complex128), ref), 1e-5)
    # a ComplexF64 slice is not on the tcgen05 path: the same call falls back to contraction + NCCL all-reduce
    A64 = random_array(np.random.default_rng(7), (24, 2 * world, 20), "complex128")
    B64 = random_array(np.random.default_rng(8), (2 * world, 20, 28), "complex128")
    ref64 = np.einsum("ihk,hkj->ij", A64, B64)
    c64, info = mdist.sharded_binary_einsum(Tensor(A64, I("ihk")), Tensor(B64, I("hkj")), I("ij"), prefer_sum=True)
    assert info[0] == _lib.SHARD_SUM and c64.on_device and len(mdist._ALLREDUCE) == 1
    check("sum_slice_c128_nccl_fallback", rel_frobenius(c64.to_host().data, ref64), 1e-12)

    kind, index, lo, hi, _ = mdist.plan_shard(Tensor(A, I(ia)), Tensor(B, I(ib)), I(ic), world, rank, prefer_sum=True)
    a_loc = mdist.local_slab(Tensor(A, I(ia)), index, lo, hi).to_device(local)
    b_loc = mdist.local_slab(Tensor(B, I(ib)), index, lo, hi).to_device(local)
    for rep in range(3):                      # repeated calls reuse the staging buffers / epochs
        slab = mdist.sum_slice_reduce_scatter(a_loc, b_loc, I(ic))
    mine = slab.data.to_host().reshape(-1, order="F")
    want = ref.reshape(-1, order="F")[rank * mine.size:(rank + 1) * mine.size]
    check("sum_slice_fused_reduce_scatter_c64", np.linalg.norm(mine - want) / np.linalg.norm(want), 1e-5)

    for mc_env in ("1", "0"):
        os.environ["MB200_DIST_MULTICAST"] = mc_env
        mdist._ALLREDUCE.clear()
        try:
            for rep in range(3):
                c_fused = mdist.sum_slice_all_reduce(a_loc, b_loc, I(ic))
            got = c_fused.to_host().data
            kinds = mdist.allreduce_plumbing_info()
            tag = f"sum_slice_fused_all_reduce_c64[{kinds[0][0]}{'+multicast' if kinds[0][1] else ''}]"
            check(tag, rel_frobenius(got.astype(np.complex128), ref), 1e-5)
            # all ranks hold the same bits
            digest = [None] * world
            import hashlib
            dist.all_gather_object(digest, hashlib.sha1(got.tobytes()).hexdigest())
            same = len(set(digest)) == 1
            ok = ok and same
            report[tag]["bit_identical_across_ranks"] = same
            if not kinds[0][1] and mc_env == "1":
                break                          # no multicast on this fabric: the second pass would repeat the first
        except Exception as e:  # noqa: BLE001
            ok = False
            report[f"sum_slice_fused_all_reduce_c64[multicast={mc_env}]"] = {"error": repr(e)[:500], "ok": False}
    torch.cuda.synchronize()
    if rank == 0:
        print("MULTI_GPU_REPORT " + json.dumps({"world": world, "ok": ok, "checks": report}), flush=True)
    dist.barrier()
    dist.destroy_process_group()
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
