"""CPU stand-in for dctdomain_b200.sharded.CudaShard built on the oracle (TEST INFRASTRUCTURE): the same five
methods, numpy / torch CPU tensors.  Used by the gloo tests to exercise the N > 1 host logic (shard bounds, bound
exchange, key exchange, merge order) without a GPU, and by the GPU tests as the independent answer."""
import numpy as np
import torch

from oracle import search_oracle as so

ID_BITS = 40
INT32_MAX = np.iinfo(np.int32).max


def pack_keys(d: np.ndarray, i: np.ndarray) -> np.ndarray:
    """faiss-style (float32 dist, int64 id, -1 padded) -> int64 bit patterns of uint64 keys dist << 40 | id."""
    keys = (d.astype(np.float64).astype(np.uint64) << np.uint64(ID_BITS)) | np.where(i >= 0, i, 0).astype(np.uint64)
    keys[i < 0] = np.uint64(0xFFFFFFFFFFFFFFFF)
    return keys.view(np.int64)


def unpack_keys(keys: np.ndarray):
    u = keys.view(np.uint64)
    empty = u == np.uint64(0xFFFFFFFFFFFFFFFF)
    d = (u >> np.uint64(ID_BITS)).astype(np.float32)
    i = (u & np.uint64((1 << ID_BITS) - 1)).astype(np.int64)
    d[empty] = so.FLT_MAX
    i[empty] = -1
    return d, i


class OracleShard:
    def __init__(self, d: int, min_bound_rows: int = 65536, threads: int = 1):
        self.d, self.rows, self.min_bound_rows, self.threads = d, np.empty((0, d), np.int8), min_bound_rows, threads
        self.calls = []

    @property
    def ntotal(self):
        return len(self.rows)

    def add(self, rows):
        self.rows = np.concatenate([self.rows, np.asarray(rows, dtype=np.int8)])

    def to_device(self, q):
        return q if isinstance(q, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(q, dtype=np.int8))

    def uses_bound(self, nq, n_min, k):
        return n_min >= self.min_bound_rows and k <= 256

    def bound(self, q, k_local, stride):
        """k_local-th smallest distance over every stride-th group of 32 rows (dctd_l1_bound's sample)."""
        self.calls.append(('bound', k_local, stride))
        n = len(self.rows)
        sel = np.concatenate([np.arange(g, min(g + 32, n)) for g in range(0, n, 32 * stride)]) if n else np.empty(0, int)
        out = np.full(len(q), INT32_MAX, dtype=np.int32)
        if len(sel) >= k_local:
            dm, _ = so.l1_topk(q.numpy(), self.rows[sel], k_local, threads=self.threads)
            out = dm[:, k_local - 1].astype(np.int32)
        return torch.from_numpy(out)

    def topk_keys(self, q, k, id_base, bound):
        self.calls.append(('topk_keys', bound is not None))
        dm, im = so.l1_topk(q.numpy(), self.rows, k, threads=self.threads)
        if bound is not None:
            drop = dm > bound.numpy()[:, None].astype(np.float32)
            dm[drop], im[drop] = so.FLT_MAX, -1            # a sorted list stays sorted: only a tail is dropped
        return torch.from_numpy(pack_keys(dm, np.where(im >= 0, im + id_base, -1)))

    def keys_merge(self, key_parts):
        parts, nq, k = key_parts.shape
        u = key_parts.numpy().view(np.uint64).transpose(1, 0, 2).reshape(nq, parts * k)
        u = np.sort(u, axis=1)[:, :k]
        d, i = unpack_keys(np.ascontiguousarray(u).view(np.int64))
        return torch.from_numpy(d), torch.from_numpy(i)
