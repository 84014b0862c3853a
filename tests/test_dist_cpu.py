"""world_size-2 gloo test of the multi-GPU plumbing on CPU: contiguous object shards, Philox streams keyed by global
object id, and the integer all-reduce of the counts (the only collective of the path)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import anytime_ref as ar


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import a3d
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    B, D, K = 11, 16, 3
    rng = np.random.Generator(np.random.PCG64(5))
    z = rng.standard_normal((B, D)).astype(np.float32)
    mask = ar.bernoulli_mask(rng, B, D, 0.5)
    mu = rng.standard_normal((6, D)).astype(np.float32)
    lo, hi = a3d.shard_range(B, rank, world)
    zc, _ = ar.impute(z[lo:hi], mask[lo:hi], mu, K, seed=3, obj_offset=lo)
    # stand-in for the per-rank GPU result: deterministic integer counts derived from the completed latents
    cnt = np.stack([(np.abs(zc).sum((1, 2)) * 1000).astype(np.int64), np.arange(lo, hi), np.ones(hi - lo, np.int64)], -1)
    total = torch.from_numpy(cnt.sum(0))
    a3d.allreduce_counts(total)
    q.put((rank, lo, hi, zc, total.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharding_matches_single_process():
    world = 2
    port = _free_port()
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    B, D, K = 11, 16, 3
    rng = np.random.Generator(np.random.PCG64(5))
    z = rng.standard_normal((B, D)).astype(np.float32)
    mask = ar.bernoulli_mask(rng, B, D, 0.5)
    mu = rng.standard_normal((6, D)).astype(np.float32)
    full, _ = ar.impute(z, mask, mu, K, seed=3)
    got = np.concatenate([r[3] for r in res], 0)
    assert np.array_equal(got, full)                  # sharding does not change any drawn sample
    cnt = np.stack([(np.abs(full).sum((1, 2)) * 1000).astype(np.int64), np.arange(B), np.ones(B, np.int64)], -1).sum(0)
    for r in res:
        assert np.array_equal(r[4], cnt)              # integer all-reduce: identical totals on every rank
    assert res[0][1] == 0 and res[0][2] == res[1][1] and res[1][2] == B
