"""Database sharded over the GPUs of one box: one process per GPU, contiguous index ranges, queries replicated,
per-rank exhaustive L1 top-k, results exchanged over NCCL / NVLink and merged by (distance, position).

Contiguous ranges keep faiss position <-> SQLite vid (reference src/query_db.py:55) and make "lower rank wins a tie"
the same as "lower position wins".  Per batch of queries (reference call site: ``index.search``, src/query_db.py:87):

  1. (large batches on large shards)  every rank bounds its ceil(k/g)-th best distance per query from a sample of its
     shard; ONE ``all_reduce(MAX)`` of nq int32 turns these into an upper bound of the k-th best distance over the
     whole database (every shard holds at least ceil(k/g) vectors within its own bound, so at least k lie within the
     largest).  A rank then reports ~k*stride/g candidates per query instead of ~k*stride: the per-rank selection work
     shrinks with the shard.
  2. per-rank scan -> the rank's k best within the bound as packed 64-bit keys (distance << 40 | global position):
     8 bytes per entry instead of faiss' 12, (distance, position) order = integer order.
  3. ONE exchange of the keys: ``all_gather`` + merge of g lists per query on every rank (results replicated), or -
     ``search_slice`` - ``all_to_all`` so that each rank merges (and keeps) the results of nq/g queries only (the
     all-vs-all workload).

The exchange is a few MB at most (3.3 MB per rank for 8192 queries, k = 50): latency- not bandwidth-bound on NVSwitch,
so the collectives are NCCL's and no fused compute+collective kernel is warranted.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib

QUERY_BATCH = 8192          # queries per exchange (bounds the candidate workspace: 64 KB per query)
STREAM_MAX_QUERIES = 16     # at most this many queries: the HBM-bound streaming regime, no bound exchange


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous range [begin, end) of rank ``rank``: r*N//g .. (r+1)*N//g."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def shared_pool(total: int, key: str, store=None):
    """Work queue over the ranks of the default process group: yields indices of ``range(total)``, every index to exactly
    one rank, in the order the ranks ask (an atomic counter ``key`` in the rendezvous store; no data-path collective).
    Fingerprinting shards by independent proteins, so a multi-GPU ``make_db`` can hand batches out this way instead of
    fixing each rank's share in advance: the ranks of a box do not get equal shares of the host's PCIe / memory bandwidth
    (measured on the 8-GPU box: 20.5 vs 35.7 GB/s per rank with all ranks copying), and with equal shares the slow ranks
    set the time.  Without a process group (or with one rank) it is ``range(total)``.  ``key`` must be new for every
    pool and the same on every rank."""
    if store is None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        store = dist.distributed_c10d._get_default_store()
    if store is None:
        yield from range(int(total))
        return
    while True:
        i = int(store.add(key, 1)) - 1
        if i >= total:
            return
        yield i


def merge_parts(dist_parts: torch.Tensor, id_parts: torch.Tensor):
    """[parts, nq, k] sorted (float32 distance, int64 id) lists -> [nq, k] (CUDA kernel dctd_l1_topk_merge)."""
    parts, nq, k = dist_parts.shape
    dev = dist_parts.device
    out_d = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    with torch.cuda.device(dev):
        rc = _lib.lib().dctd_l1_topk_merge(dist_parts.contiguous().data_ptr(), id_parts.contiguous().data_ptr(),
                                           parts, nq, k, out_d.data_ptr(), out_i.data_ptr(),
                                           torch.cuda.current_stream(dev).cuda_stream)
    _lib.check(rc, 'dctd_l1_topk_merge')
    return out_d, out_i


class CudaShard:
    """This rank's shard on its GPU: the four device steps of a sharded search (libdctd kernels)."""

    def __init__(self, d: int, device=None):
        from . import index as dindex
        self._dindex = dindex
        self.index = dindex.IndexFlatL1(d, device)

    @property
    def ntotal(self):
        return self.index.ntotal

    def add(self, rows):
        self.index.add(rows)

    def to_device(self, q):
        if isinstance(q, torch.Tensor):
            return q.to(self.index._dev)
        return torch.from_numpy(self._dindex._as_int8(q, self.index.d)).to(self.index._dev)

    def uses_bound(self, nq, n_min, k):
        return self.index.uses_bound(nq, n_min, k)

    def bound(self, q, k_local, stride):
        return self.index.bound_device(q, k_local, stride)

    def topk_keys(self, q, k, id_base, bound):
        return self.index.search_keys_device(q, k, id_base=id_base, bound=bound)

    def keys_merge(self, key_parts):
        return self._dindex.keys_merge(key_parts)


class ShardedIndex:
    """This rank's shard plus the exchange.  ``shard`` defaults to ``CudaShard`` (the CUDA kernels); the CPU (gloo)
    tests inject a stand-in with the same five methods to exercise the partitioning and the collectives without a GPU."""

    def __init__(self, d: int, n_total: int, rank: int = None, world: int = None, group=None, shard=None, device=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.d, self.n_total = d, n_total
        self.begin, self.end = shard_bounds(n_total, self.world, self.rank)
        self.n_min = min(shard_bounds(n_total, self.world, r)[1] - shard_bounds(n_total, self.world, r)[0]
                         for r in range(self.world))
        self.shard = shard if shard is not None else CudaShard(d, device)
        self.index = getattr(self.shard, 'index', None)
        self.last_path = None           # for tests / the bench line: which exchange the last batch took
        self.events = None              # set to a list to have a CUDA event recorded after every step (bench phases)

    def _mark(self, name):
        if self.events is not None:
            ev = torch.cuda.Event(enable_timing=True)
            ev.record()
            self.events.append((name, ev))

    def add_local(self, rows):
        """Rows [begin, end) of the global database (int8 [end-begin, d])."""
        if rows.shape[0] != self.end - self.begin:
            raise ValueError(f'rank {self.rank} expects {self.end - self.begin} rows, got {rows.shape[0]}')
        self.shard.add(rows)

    # ---------------------------------------------------------------------------------------------------------
    def sample_stride(self):
        """Sampling stride of the bound step: candidates per rank ~ 1.5 k stride / g, sample + rescan cost ~ 2 / stride
        of the scan."""
        return 64 if self.world >= 8 else 32

    def _local_keys(self, q, k):
        """Steps 1 + 2 for one batch: (keys int64 [nq, k], took the bound exchange?)."""
        nq = int(q.shape[0])
        bound = None
        self._mark('start')
        if self.world > 1 and nq > STREAM_MAX_QUERIES and self.shard.uses_bound(nq, self.n_min, k):
            k_local = -(-k // self.world)
            bound = self.shard.bound(q, k_local, self.sample_stride())
            self._mark('bound')
            dist.all_reduce(bound, op=dist.ReduceOp.MAX, group=self.group)
            self._mark('all_reduce')
        keys = self.shard.topk_keys(q, k, self.begin, bound)
        self._mark('scan_select')
        return keys, bound is not None

    def search(self, q, k: int):
        """Replicated queries [nq, d] (int8 CUDA tensor, or any host array as ``faiss`` ``index.search`` takes) ->
        (float32 [nq, k], int64 [nq, k]) identical on every rank; CUDA tensors for a CUDA input, numpy arrays otherwise."""
        host = not (isinstance(q, torch.Tensor) and q.is_cuda)
        qd = self.shard.to_device(q) if host else q
        nq = int(qd.shape[0])
        if nq == 0:
            dm = torch.empty((0, k), dtype=torch.float32, device=qd.device)
            im = torch.empty((0, k), dtype=torch.int64, device=qd.device)
            return (dm.cpu().numpy(), im.cpu().numpy()) if host else (dm, im)
        out_d, out_i = [], []
        for b in range(0, nq, QUERY_BATCH):
            qb = qd[b:b + QUERY_BATCH]
            keys, bounded = self._local_keys(qb, k)
            n = int(qb.shape[0])
            if self.world == 1:
                parts = keys.view(1, n, k)
                self.last_path = 'local'
            else:
                allk = torch.empty((self.world * n, k), dtype=keys.dtype, device=keys.device)
                dist.all_gather_into_tensor(allk, keys.contiguous(), group=self.group)      # rank-major blocks
                parts = allk.view(self.world, n, k)
                self.last_path = 'bound+all_gather' if bounded else 'all_gather'
                self._mark('exchange')
            d_, i_ = self.shard.keys_merge(parts)
            self._mark('merge')
            out_d.append(d_)
            out_i.append(i_)
        dm = out_d[0] if len(out_d) == 1 else torch.cat(out_d)
        im = out_i[0] if len(out_i) == 1 else torch.cat(out_i)
        if host:
            return dm.cpu().numpy(), im.cpu().numpy()
        return dm, im

    def search_slice(self, q, k: int):
        """The all-vs-all form: every rank scans its shard for ALL queries, but merges and keeps the results of its
        slice of the queries only (``all_to_all`` of the keys instead of ``all_gather``).  Returns (dist, ids, q_begin,
        q_end): rows [q_begin, q_end) of the full result, q_begin = rank * ceil(nq / g) (clipped to nq)."""
        host = not (isinstance(q, torch.Tensor) and q.is_cuda)
        qd = self.shard.to_device(q) if host else q
        nq = int(qd.shape[0])
        if nq == 0 or nq > QUERY_BATCH:
            raise ValueError(f'search_slice takes 1..{QUERY_BATCH} queries per call')
        g = self.world
        per = -(-nq // g)
        if per * g != nq:                                   # equal splits for the collective: repeat the last query
            pad = qd[-1:].expand(per * g - nq, qd.shape[1])
            qd = torch.cat([qd, pad])
        keys, bounded = self._local_keys(qd, k)
        if g == 1:
            recv = keys.view(1, per, k)
            self.last_path = 'local'
        else:
            recv = torch.empty_like(keys)
            dist.all_to_all_single(recv, keys.contiguous(), group=self.group)               # block r <- rank r's keys
            recv = recv.view(g, per, k)
            self.last_path = 'bound+all_to_all' if bounded else 'all_to_all'
            self._mark('exchange')
        dm, im = self.shard.keys_merge(recv)
        self._mark('merge')
        qb, qe = min(nq, self.rank * per), min(nq, (self.rank + 1) * per)
        dm, im = dm[:qe - qb], im[:qe - qb]
        if host:
            return dm.cpu().numpy(), im.cpu().numpy(), qb, qe
        return dm, im, qb, qe
