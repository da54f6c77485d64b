"""The sharded search on real GPUs, compared bit for bit with the oracle over the WHOLE database
(replaces the reference's index.search at src/query_db.py:75-76,87, position <-> vid contract of :55).

  * test_sharded_pieces_one_gpu   - every CUDA step of the N > 1 path (dctd_l1_bound, dctd_l1_topk_keys with an
                                    external bound, dctd_l1_keys_merge) with the ranks simulated on one device and the
                                    MAX reduction done in torch: runs on the driver's 1-GPU box.
  * test_sharded_nccl_*           - the real thing: one process per GPU, NCCL all_reduce / all_gather / all_to_all.
                                    Needs >= 2 GPUs (skipped otherwise); `bench.py` repeats the comparison inside the
                                    driver's multi-GPU runs (search.parity).
"""
import os
import socket

import numpy as np
import pytest
import torch

import synth

pytestmark = pytest.mark.gpu
K = 50


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    port = s.getsockname()[1]
    s.close()
    return port


def make_db(world: int, per_shard: int, seed: int = 11):
    """Synthetic fingerprints with duplicates straddling EVERY shard boundary (ties broken by position across ranks)."""
    from dctdomain_b200.sharded import shard_bounds
    n = per_shard * world + 3                       # ragged: the shards differ in size
    db = synth.fingerprints(seed, n)
    for r in range(1, world):
        b = shard_bounds(n, world, r)[0]
        db[b - 12:b + 12] = db[5 + r]               # 24 equal rows across the boundary, equal to an early row
    db[n - 9:] = db[3]                              # ... and at the very end of the last shard
    return db


def make_queries(db, world, nq, seed=5):
    """Rows of the database (so that the planted duplicates are everybody's nearest neighbours) + random perturbations."""
    rs = np.random.RandomState(seed)
    rows = np.concatenate([np.arange(3, 6 + world), rs.randint(0, len(db), size=max(0, nq - 3 - world))])[:nq]
    q = db[rows].astype(np.int16)
    q[len(q) // 2:] += rs.randint(-2, 3, size=q[len(q) // 2:].shape).astype(np.int16)
    return np.clip(q, 0, 127).astype(np.int8)


def test_sharded_pieces_one_gpu():
    from dctdomain_b200.sharded import CudaShard, shard_bounds
    from oracle import search_oracle as so
    world, per = 4, 70_000
    db = make_db(world, per)
    q = make_queries(db, world, 40)
    dm, im = so.l1_topk(q, db, K, threads=4)
    shards = []
    for r in range(world):
        b, e = shard_bounds(len(db), world, r)
        s = CudaShard(480)
        s.add(db[b:e])
        shards.append((s, b))
    qd = shards[0][0].to_device(q)
    assert shards[0][0].uses_bound(len(q), per, K)
    k_local = -(-K // world)
    for stride in (32, 64):
        bounds = torch.stack([s.bound(qd, k_local, stride) for s, _ in shards])
        bound = bounds.max(dim=0).values                                        # = all_reduce(MAX)
        assert int(bound.max()) < 2 ** 31 - 1
        keys = torch.stack([s.topk_keys(qd, K, b, bound) for s, b in shards])
        d, i = shards[0][0].keys_merge(keys)
        assert np.array_equal(i.cpu().numpy(), im) and np.array_equal(d.cpu().numpy(), dm)
    # without a bound, and with the streaming regime (<= 16 queries)
    for nq in (40, 7):
        keys = torch.stack([s.topk_keys(qd[:nq], K, b, None) for s, b in shards])
        d, i = shards[0][0].keys_merge(keys)
        assert np.array_equal(i.cpu().numpy(), im[:nq]) and np.array_equal(d.cpu().numpy(), dm[:nq])
    # a bound that is NOT an upper bound of the k-th distance truncates the lists but never invents or reorders entries
    tight = torch.full((len(q),), 0, dtype=torch.int32, device=qd.device)
    keys = shards[0][0].topk_keys(qd, K, 0, tight)
    d0, i0 = shards[0][0].keys_merge(keys.view(1, len(q), K))
    b0, e0 = shard_bounds(len(db), world, 0)
    dl, il = so.l1_topk(q, db[b0:e0], K, threads=4)
    want_i = np.where(dl == 0, il, -1)
    assert np.array_equal(i0.cpu().numpy(), want_i)


def _worker(rank, world, port, out_dir, per_shard, nq_list):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, 'tests')]
    import torch.distributed as dist
    from dctdomain_b200.sharded import ShardedIndex
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group('nccl', rank=rank, world_size=world, device_id=torch.device('cuda', rank))
    db = make_db(world, per_shard)
    sh = ShardedIndex(480, len(db))
    sh.add_local(db[sh.begin:sh.end])
    out = {}
    for nq in nq_list:
        q = make_queries(db, world, nq)
        d, i = sh.search(torch.from_numpy(q).cuda(), K)                  # device tensors
        out[f'dev{nq}_d'], out[f'dev{nq}_i'], out[f'dev{nq}_path'] = d.cpu().numpy(), i.cpu().numpy(), sh.last_path
        d, i = sh.search(q.astype(np.float32), K)                        # faiss-style host arrays (any numeric dtype)
        out[f'host{nq}_d'], out[f'host{nq}_i'] = d, i
        d, i, qb, qe = sh.search_slice(torch.from_numpy(q).cuda(), K)
        out[f'sl{nq}_d'], out[f'sl{nq}_i'], out[f'sl{nq}_b'], out[f'sl{nq}_e'] = d.cpu().numpy(), i.cpu().numpy(), qb, qe
        out[f'sl{nq}_path'] = sh.last_path
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), **out)
    dist.barrier()
    dist.destroy_process_group()


def _run(tmp_path, world, per_shard, nq_list, expect_bound):
    import torch.multiprocessing as mp
    from oracle import search_oracle as so
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), per_shard, nq_list), nprocs=world, join=True)
    db = make_db(world, per_shard)
    for nq in nq_list:
        q = make_queries(db, world, nq)
        dm, im = so.l1_topk(q, db, K, threads=8)
        covered = np.zeros(nq, bool)
        for r in range(world):
            z = np.load(tmp_path / f'r{r}.npz')
            for kind in ('dev', 'host'):
                assert np.array_equal(z[f'{kind}{nq}_i'], im), (kind, nq, r)
                assert np.array_equal(z[f'{kind}{nq}_d'], dm), (kind, nq, r)
            bounded = expect_bound and nq > 16
            assert str(z[f'dev{nq}_path']) == ('bound+all_gather' if bounded else 'all_gather')
            assert str(z[f'sl{nq}_path']) == ('bound+all_to_all' if bounded else 'all_to_all')
            qb, qe = int(z[f'sl{nq}_b']), int(z[f'sl{nq}_e'])
            assert np.array_equal(z[f'sl{nq}_i'], im[qb:qe]) and np.array_equal(z[f'sl{nq}_d'], dm[qb:qe])
            covered[qb:qe] = True
        assert covered.all()


def _need(n):
    if torch.cuda.device_count() < n:
        pytest.skip(f'needs {n} GPUs')


def test_sharded_nccl_two_ranks_large_shards(tmp_path):
    """200k-row class database, shards on the threshold path: bound all-reduce + key all-gather / all-to-all."""
    _need(2)
    _run(tmp_path, 2, 100_000, [37, 7], expect_bound=True)


def test_sharded_nccl_two_ranks_shard_smaller_than_k(tmp_path):
    """Shards with fewer than k rows (heap path, padding inside the per-rank lists); n_total both above and below k."""
    _need(2)
    for sub, per in (('a', 30), ('b', 11)):
        (tmp_path / sub).mkdir()
        _run(tmp_path / sub, 2, per, [21, 3], expect_bound=False)


def test_sharded_nccl_all_gpus(tmp_path):
    """Every GPU of the box (8 on the driver's scale box), shards on the threshold path."""
    n = torch.cuda.device_count()
    _need(3)
    _run(tmp_path, n, 66_000, [45, 8], expect_bound=True)
