"""CPU, world_size 2, gloo: the multi-GPU host logic (shard planning, slab distribution, assembling by
all_gather / all_reduce). The per-shard contraction is injected (the oracle) because the product's
contraction needs a B200; on the GPU box the same code path runs with BackendB200 + NCCL (bench.py --gpus N)."""
import os
import socket
import sys

import numpy as np
import pytest
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import muscle_b200 as mb
        from muscle_b200 import Index, Tensor, _lib
        from muscle_b200.dist import sharded_binary_einsum
        from oracle import binary_einsum_general, rel_frobenius
        from cases import random_array

        I = lambda s: [Index(c) for c in s]

        def oracle_contract(inds_c, a, b):
            tags = lambda t: [i.tag for i in t.inds]
            c = binary_einsum_general([i.tag for i in inds_c], a.data, tags(a), b.data, tags(b))
            return Tensor(c, inds_c)

        rng = np.random.default_rng(42)           # same seed on every rank: replicated operands
        res = {}
        # (1) free-index shard (config-4 pattern): no collective, slabs assembled only for the check
        a = random_array(rng, (4, 3, 5, 4, 3, 5), "complex128")
        b = random_array(rng, (5, 2, 3, 3, 4, 6), "complex128")
        ref = binary_einsum_general(list("abcghi"), a, list("adbecf"), b, list("fgdhei"))
        c, info = sharded_binary_einsum(Tensor(a, I("adbecf")), Tensor(b, I("fgdhei")), I("abcghi"),
                                        gather=True, contract=oracle_contract)
        res["free_kind"] = info[0] == _lib.SHARD_FREE and info[1] == Index("i")
        res["free_err"] = rel_frobenius(c.data, ref)
        c_slab, info = sharded_binary_einsum(Tensor(a, I("adbecf")), Tensor(b, I("fgdhei")), I("abcghi"),
                                             contract=oracle_contract)
        res["free_slab_shape"] = c_slab.shape == (4, 5, 3, 2, 3, 3)
        res["free_slab_err"] = rel_frobenius(c_slab.data, ref[..., info[2]:info[3]])
        # (2) summed-index slice + all_reduce(SUM) (config-5 pattern)
        c, info = sharded_binary_einsum(Tensor(a, I("adbecf")), Tensor(b, I("fgdhei")), I("abcghi"),
                                        prefer_sum=True, contract=oracle_contract)
        res["sum_kind"] = info[0] == _lib.SHARD_SUM
        res["sum_err"] = rel_frobenius(c.data, ref)
        # (3) batch-index shard
        a3 = random_array(rng, (6, 5, 4), "complex64")
        b3 = random_array(rng, (5, 7, 4), "complex64")
        ref3 = binary_einsum_general(list("kiz"), a3, list("ijz"), b3, list("jkz"))
        c, info = sharded_binary_einsum(Tensor(a3, I("ijz")), Tensor(b3, I("jkz")), I("kiz"), gather=True,
                                        contract=oracle_contract)
        res["batch_kind"] = info[0] == _lib.SHARD_BATCH
        res["batch_err"] = rel_frobenius(c.data, ref3)
        # (4) nothing to shard → replicas only
        c, info = sharded_binary_einsum(Tensor(3 * np.ones((1, 1)), I("ij")), Tensor(np.ones((1, 1)), I("jk")), I("ik"),
                                        contract=oracle_contract)
        res["none_kind"] = info[0] == _lib.SHARD_NONE and float(c.data[0, 0]) == 3.0
        q.put((rank, res))
    except Exception as e:  # surface the failure instead of letting the parent wait for the timeout
        q.put((rank, {"error": repr(e)}))
        raise
    finally:
        dist.destroy_process_group()


def test_sharded_binary_einsum_world2_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = [q.get(timeout=180) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, res in results:
        assert "error" not in res, res
        assert res["free_kind"] and res["free_slab_shape"] and res["sum_kind"] and res["batch_kind"] and res["none_kind"], res
        assert res["free_err"] < 1e-13 and res["free_slab_err"] < 1e-13 and res["sum_err"] < 1e-13, res
        assert res["batch_err"] < 1e-5, res
