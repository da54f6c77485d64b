"""Database sharded over the GPUs of one box: one process per GPU, contiguous index ranges, queries
replicated, per-rank exhaustive L1 top-k, ONE collective (all-gather of the [nq, k] results over
NCCL / NVLink), k-way merge by (distance, position) on every rank.

Contiguous ranges keep faiss position <-> SQLite vid (reference src/query_db.py:55) and make
"lower rank wins a tie" the same as "lower position wins".  The exchange is 12 bytes per (query, k)
entry per rank (48 MB for 10k queries, k = 50, 8 ranks), so it is latency- not bandwidth-bound on
NVSwitch; no fused compute+collective kernel is warranted here.
"""
from __future__ import annotations

import torch
import torch.distributed as dist

from . import _lib


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous range [begin, end) of rank ``rank``: r*N//g .. (r+1)*N//g."""
    return (n_total * rank) // world, (n_total * (rank + 1)) // world


def merge_parts(dist_parts: torch.Tensor, id_parts: torch.Tensor):
    """[parts, nq, k] sorted lists -> [nq, k] (CUDA kernel dctd_l1_topk_merge)."""
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


class ShardedIndex:
    """This rank's shard plus the exchange.  ``local_search(q, k, id_base)`` and ``merge(dist_parts,
    id_parts)`` default to the CUDA kernels; the CPU (gloo) tests inject stand-ins to exercise the
    partitioning and the collective without a GPU."""

    def __init__(self, d: int, n_total: int, rank: int = None, world: int = None, group=None,
                 local_search=None, merge=None, device=None):
        self.group = group
        self.rank = dist.get_rank(group) if rank is None else rank
        self.world = dist.get_world_size(group) if world is None else world
        self.d, self.n_total = d, n_total
        self.begin, self.end = shard_bounds(n_total, self.world, self.rank)
        self._local_search = local_search
        self._merge = merge or merge_parts
        self.index = None
        if local_search is None:
            from . import index as dindex
            self.index = dindex.IndexFlatL1(d, device)

    def add_local(self, rows):
        """Rows [begin, end) of the global database (int8 [end-begin, d])."""
        if rows.shape[0] != self.end - self.begin:
            raise ValueError(f'rank {self.rank} expects {self.end - self.begin} rows, got {rows.shape[0]}')
        self.index.add(rows)

    def search(self, q: torch.Tensor, k: int):
        """Replicated int8 queries [nq, d] -> (float32 [nq, k], int64 [nq, k]) identical on every rank."""
        if self._local_search is not None:
            d_loc, i_loc = self._local_search(q, k, self.begin)
        else:
            d_loc, i_loc = self.index.search_device(q, k, id_base=self.begin)
        if self.world == 1:
            return d_loc, i_loc
        nq, kk = d_loc.shape
        d_all = torch.empty((self.world * nq, kk), dtype=d_loc.dtype, device=d_loc.device)
        i_all = torch.empty((self.world * nq, kk), dtype=i_loc.dtype, device=i_loc.device)
        dist.all_gather_into_tensor(d_all, d_loc.contiguous(), group=self.group)   # rank-major blocks
        dist.all_gather_into_tensor(i_all, i_loc.contiguous(), group=self.group)
        return self._merge(d_all.view(self.world, nq, kk), i_all.view(self.world, nq, kk))
