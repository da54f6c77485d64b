"""Seeded synthetic ESM-2-shaped inputs shared by the golden generator, the tests and bench.py.

No ESM-2 weights are available offline, so embeddings are synthetic:
  white   N(0,1)
  esm     per-dimension offset + 0.15-sigma random walk along L + 0.3-sigma noise + 8 outlier dims x20
  offset  5 + 0.01*N(0,1)   (large common offset: the float32 stress case)
``np.random.RandomState`` is used because its streams are frozen across numpy versions.
"""
from __future__ import annotations

import numpy as np

KINDS = ('white', 'esm', 'offset')


def embedding(seed: int, n_rows: int, dim: int, kind: str = 'white') -> np.ndarray:
    rs = np.random.RandomState(seed)
    if kind == 'white':
        x = rs.standard_normal((n_rows, dim))
    elif kind == 'esm':
        off = rs.standard_normal(dim) * 0.5
        walk = np.cumsum(rs.standard_normal((n_rows, dim)) * 0.15, axis=0)
        x = off[None, :] + walk + rs.standard_normal((n_rows, dim)) * 0.3
        hot = rs.choice(dim, size=min(8, dim), replace=False)
        x[:, hot] *= 20.0
    elif kind == 'offset':
        x = 5.0 + 0.01 * rs.standard_normal((n_rows, dim))
    else:
        raise ValueError(kind)
    return np.ascontiguousarray(x, dtype=np.float32)


def layers(seed: int, n_rows: int, dim: int, kind: str = 'white', ids=(15, 21)) -> dict:
    """{layer_id: float32[n_rows, dim]} in the reference's insertion order (embedding.py:174-175)."""
    return {lid: embedding(seed * 1000 + 7 * i + 1, n_rows, dim, kind) for i, lid in enumerate(ids)}


def fingerprints(seed: int, n: int, dim: int = 480) -> np.ndarray:
    """int8 rows shaped like real fingerprints: per group of 80 bytes one 0 and one 127,
    the rest clip(N(63.6, 27.7)) (fixture statistics)."""
    rs = np.random.RandomState(seed)
    x = np.clip(np.rint(rs.normal(63.6, 27.7, size=(n, dim))), 1, 126).astype(np.int8)
    for g in range(0, dim - dim % 80, 80):
        lo = rs.randint(0, 80, size=n)
        hi = (lo + 1 + rs.randint(0, 79, size=n)) % 80
        x[np.arange(n), g + lo] = 0
        x[np.arange(n), g + hi] = 127
    return x


def random_partition(rs: np.random.RandomState, length: int, pieces: int, min_len: int = 3):
    """Domain strings partitioning 1..length into ``pieces`` contiguous parts."""
    pieces = max(1, min(pieces, length // min_len))
    cuts = sorted(rs.choice(np.arange(1, length // min_len), size=pieces - 1, replace=False) * min_len) \
        if pieces > 1 else []
    edges = [0] + [int(c) for c in cuts] + [length]
    return [f'{a + 1}-{b}' for a, b in zip(edges[:-1], edges[1:])]
