"""CPU oracle for hot path 2 (L1 top-k search + dct-sim).  TEST INFRASTRUCTURE - see oracle/__init__.py.

  * ``l1_topk``            - ctypes wrapper over oracle/l1_flat.c (faiss 1.7.4 flat-L1 restated);
  * ``l1_topk_numpy``      - independent brute force (integer distances, lexsort by (dist, id)),
                             for small cases and to cross-check the C code;
  * ``top_hits_lines``     - the ordering / cut / score rules of ``get_top_hits``
                             (reference src/query_db.py:17-59) producing its log lines;
  * ``prost_similarity`` / ``domain_sim`` - reference src/dct-sim.py:12-50.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None
FLT_MAX = np.finfo(np.float32).max


def build(force: bool = False) -> str:
    """Compile oracle/l1_flat.c (make); returns the library path."""
    path = os.path.join(_HERE, '_build', 'libl1oracle.so')
    if force or not os.path.exists(path):
        subprocess.run(['make', '-C', _HERE], check=True, stdout=subprocess.DEVNULL)
    return path


def _lib():
    global _LIB
    if _LIB is None:
        lib = ctypes.CDLL(build())
        for name, qt in (('oracle_l1_topk_i8', ctypes.c_void_p), ('oracle_l1_topk_f32', ctypes.c_void_p)):
            fn = getattr(lib, name)
            fn.restype = ctypes.c_int
            fn.argtypes = [qt, ctypes.c_int64, ctypes.c_void_p, ctypes.c_int64, ctypes.c_int,
                           ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _LIB = lib
    return _LIB


def l1_topk(q: np.ndarray, db: np.ndarray, k: int, threads: int = 1):
    """(dist float32 [nq,k], ids int64 [nq,k]) - faiss flat-L1 semantics."""
    q = np.ascontiguousarray(q)
    db = np.ascontiguousarray(db)
    nq, d = q.shape
    nb = db.shape[0]
    out_d = np.empty((nq, k), dtype=np.float32)
    out_i = np.empty((nq, k), dtype=np.int64)
    if q.dtype == np.int8 and db.dtype == np.int8:
        fn = _lib().oracle_l1_topk_i8
    else:
        q = np.ascontiguousarray(q, dtype=np.float32)
        db = np.ascontiguousarray(db, dtype=np.float32)
        fn = _lib().oracle_l1_topk_f32
    rc = fn(q.ctypes.data, nq, db.ctypes.data, nb, d, k, out_d.ctypes.data, out_i.ctypes.data, threads)
    if rc != 0:
        raise RuntimeError(f'oracle_l1_topk failed: {rc}')
    return out_d, out_i


def l1_topk_numpy(q: np.ndarray, db: np.ndarray, k: int):
    """Brute force: integer |q-x| sums, k smallest by (dist, id), (-1, FLT_MAX) padding."""
    q = np.asarray(q).astype(np.int64)
    db = np.asarray(db).astype(np.int64)
    nq, nb = q.shape[0], db.shape[0]
    out_d = np.full((nq, k), FLT_MAX, dtype=np.float32)
    out_i = np.full((nq, k), -1, dtype=np.int64)
    for i in range(nq):
        dist = np.abs(db - q[i]).sum(axis=1)
        order = np.lexsort((np.arange(nb), dist))[:k]
        out_d[i, :len(order)] = dist[order].astype(np.float32)
        out_i[i, :len(order)] = order
    return out_d, out_i


def top_hits_lines(dm, im, top, q_labels, db_labels):
    """Log lines of get_top_hits for metric 'l1' (reference src/query_db.py:33-59).

    ``q_labels[i]`` = (pid, domain) of query row i; ``db_labels[j]`` = (pid, domain) of
    database position j (faiss position j <-> SQLite vid j+1, query_db.py:55).
    """
    flat = {}
    for i, row in enumerate(dm):
        for j, dist in enumerate(row):
            flat[i, j] = dist
    flat = dict(sorted(flat.items(), key=lambda kv: kv[1]))   # stable: ties keep (i, j) order
    lines = []
    for rank, (i, j) in enumerate(list(flat.keys())[:top]):
        pos = int(im[i, j])
        if pos == -1:
            break
        score = round(1 - (flat[i, j] / 17000), 4)
        q_pid, q_dom = q_labels[i]
        d_pid, d_dom = db_labels[pos]
        lines.append('Query: %s %s, Result %s: %s %s, Similarity: %s'
                     % (q_pid, q_dom, rank + 1, d_pid, d_dom, score))
    return lines


def prost_similarity(a: np.ndarray, b: np.ndarray) -> float:
    """reference src/dct-sim.py:12-26."""
    d = abs(a - b).sum()
    d /= 17000
    d = min(d, 1)
    return 1 - d


def domain_sim(fi: np.ndarray, fj: np.ndarray):
    """reference src/dct-sim.py:28-50: (max over all pairs, similarity of the last pair)."""
    best = 0
    s = None
    for a in range(fi.shape[0]):
        for b in range(fj.shape[0]):
            s = prost_similarity(fi[a], fj[b])
            if s > best:
                best = s
    return best, s
