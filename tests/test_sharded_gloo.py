"""world_size-2 gloo test (CPU) of the sharded search plumbing: contiguous shard bounds, id_base, the
all-gather layout and the merge order.  The per-rank search and the merge kernel are CUDA in the product;
here the oracle stands in for them (injected), so only the host-side N>1 logic is under test."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import synth


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, 'tests')]
    from dctdomain_b200.sharded import ShardedIndex
    from oracle import search_oracle as so
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    db = synth.fingerprints(4, 3001)
    db[1500:1520] = db[2]          # ties that straddle the shard boundary
    q = db[:25]
    k = 50

    def local_search(qt, kk, id_base):
        b, e = sh.begin, sh.end
        d, i = so.l1_topk(qt.numpy(), db[b:e], kk)
        i = np.where(i >= 0, i + id_base, -1)
        return torch.from_numpy(d), torch.from_numpy(i)

    def merge(dp, ip):
        parts, nq, kk = dp.shape
        d = dp.permute(1, 0, 2).reshape(nq, parts * kk).numpy()
        i = ip.permute(1, 0, 2).reshape(nq, parts * kk).numpy()
        od, oi = np.empty((nq, kk), np.float32), np.empty((nq, kk), np.int64)
        for r in range(nq):
            key = np.lexsort((np.where(i[r] < 0, np.iinfo(np.int64).max, i[r]), d[r]))[:kk]
            od[r], oi[r] = d[r][key], i[r][key]
        return torch.from_numpy(od), torch.from_numpy(oi)

    sh = ShardedIndex(480, len(db), local_search=local_search, merge=merge)
    assert (sh.begin, sh.end) == ((0, 1500) if rank == 0 else (1500, 3001))
    d, i = sh.search(torch.from_numpy(q), k)
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), d=d.numpy(), i=i.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_search_equals_single(tmp_path):
    from oracle import search_oracle as so
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    db = synth.fingerprints(4, 3001)
    db[1500:1520] = db[2]
    dm, im = so.l1_topk(db[:25], db, 50)
    for r in range(world):
        z = np.load(tmp_path / f'r{r}.npz')
        assert np.array_equal(z['i'], im) and np.array_equal(z['d'], dm)


def test_shard_bounds_cover_everything():
    from dctdomain_b200.sharded import shard_bounds
    for n in (0, 1, 7, 43, 1_000_000, 50_000_000):
        for g in (1, 2, 4, 8):
            b = [shard_bounds(n, g, r) for r in range(g)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(g - 1))
