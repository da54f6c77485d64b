"""world_size-2/3 gloo tests (CPU) of the sharded search plumbing: contiguous shard bounds, id_base, the bound
all-reduce, the all-gather / all-to-all layouts and the merge order.  The per-rank steps are CUDA in the product;
here the oracle stands in for them (tests/oracle_shard.py, injected), so only the host-side N>1 logic is under
test.  The CUDA + NCCL path itself is compared with the oracle in tests/test_sharded_nccl_gpu.py."""
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


def _database():
    db = synth.fingerprints(4, 3001)
    db[1500:1520] = db[2]          # ties that straddle the shard boundary
    db[2990:3001] = db[7]          # ... and sit at the very end of the last shard
    return db


def _worker(rank, world, port, out_dir, min_bound_rows):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root, os.path.join(root, 'tests')]
    from dctdomain_b200.sharded import ShardedIndex, shard_bounds
    from oracle_shard import OracleShard
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    db = _database()
    k = 50
    sh = ShardedIndex(480, len(db), shard=OracleShard(480, min_bound_rows=min_bound_rows))
    assert (sh.begin, sh.end) == shard_bounds(len(db), world, rank)
    sh.add_local(db[sh.begin:sh.end])
    out = {}
    for name, nq in (('few', 7), ('many', 45)):
        d, i = sh.search(db[:nq], k)                       # host arrays in, host arrays out
        out[name + '_d'], out[name + '_i'], out[name + '_path'] = d, i, sh.last_path
    d, i, qb, qe = sh.search_slice(torch.from_numpy(db[:45]), k)
    out.update(slice_d=d, slice_i=i, slice_b=qb, slice_e=qe, slice_path=sh.last_path)
    np.savez(os.path.join(out_dir, f'r{rank}.npz'), **out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world,min_bound_rows', [(2, 65536), (2, 512), (3, 512)])
def test_sharded_search_equals_single(tmp_path, world, min_bound_rows):
    """all_gather path, bound exchange (forced on for tiny shards by min_bound_rows=512) and all_to_all slices against
    the oracle over the whole database."""
    from oracle import search_oracle as so
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path), min_bound_rows), nprocs=world, join=True)
    db = _database()
    dm, im = so.l1_topk(db[:45], db, 50)
    bounded = min_bound_rows <= len(db) // world
    covered = np.zeros(45, bool)
    for r in range(world):
        z = np.load(tmp_path / f'r{r}.npz')
        assert np.array_equal(z['few_i'], im[:7]) and np.array_equal(z['few_d'], dm[:7])
        assert np.array_equal(z['many_i'], im) and np.array_equal(z['many_d'], dm)
        assert str(z['few_path']) == 'all_gather'                                   # <= 16 queries: no bound exchange
        assert str(z['many_path']) == ('bound+all_gather' if bounded else 'all_gather')
        assert str(z['slice_path']) == ('bound+all_to_all' if bounded else 'all_to_all')
        qb, qe = int(z['slice_b']), int(z['slice_e'])
        assert np.array_equal(z['slice_i'], im[qb:qe]) and np.array_equal(z['slice_d'], dm[qb:qe])
        covered[qb:qe] = True
    assert covered.all()


def test_shard_bounds_cover_everything():
    from dctdomain_b200.sharded import shard_bounds
    for n in (0, 1, 7, 43, 1_000_000, 50_000_000):
        for g in (1, 2, 4, 8):
            b = [shard_bounds(n, g, r) for r in range(g)]
            assert b[0][0] == 0 and b[-1][1] == n and all(b[i][1] == b[i + 1][0] for i in range(g - 1))


def _pool_worker(rank, world, port, out_dir):
    import sys
    import time
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    sys.path[:0] = [root]
    from dctdomain_b200.sharded import shared_pool
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    got = []
    for i in shared_pool(40, 'pool_a'):
        got.append(i)
        time.sleep(0.002 * (1 + 3 * rank))          # rank 0 is the fast one
    second = list(shared_pool(5, 'pool_b'))          # a fresh key starts from 0 again
    empty = list(shared_pool(0, 'pool_c'))
    np.savez(os.path.join(out_dir, f'p{rank}.npz'), got=np.array(got, dtype=np.int64), second=np.array(second, dtype=np.int64),
             empty=np.array(empty, dtype=np.int64))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize('world', [2, 3])
def test_shared_pool_hands_every_index_to_exactly_one_rank(tmp_path, world):
    """The work queue bench.py's multi-rank e2e uses (and a multi-GPU make_db would): every index once over all ranks,
    ascending within a rank, the faster rank takes more; without a process group it is range(total)."""
    from dctdomain_b200.sharded import shared_pool
    assert list(shared_pool(7, 'no_group')) == list(range(7))
    mp.spawn(_pool_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    parts = [np.load(tmp_path / f'p{r}.npz') for r in range(world)]
    allgot = np.concatenate([z['got'] for z in parts])
    assert sorted(allgot.tolist()) == list(range(40))
    for z in parts:
        assert np.all(np.diff(z['got']) > 0) and len(z['empty']) == 0
    assert len(parts[0]['got']) > len(parts[-1]['got'])
    assert sorted(np.concatenate([z['second'] for z in parts]).tolist()) == list(range(5))
