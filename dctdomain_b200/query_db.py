"""Drop-in for the search half of reference ``src/query_db.py`` (get_top_hits :17-59, search_db :62-91).

``search_db`` keeps the reference's signature and output (one ``logging.info`` line per hit, same
text), but gathers the fingerprints of every query protein into ONE exhaustive L1 top-k launch on
the GPU (dctdomain_b200.index) and regroups per protein on the host, instead of one tiny
``index.search`` per protein (nq = 1..13).  SQLite access stays as in the reference: the two
database arguments are opened with the reference's ``Database`` class (or any object with ``cur``,
``load_fprints(pid)`` and ``close()``).
"""
from __future__ import annotations

import argparse
import logging

import numpy as np

from . import index as dindex


def get_top_hits(dm: np.ndarray, im: np.ndarray, top: int, fp_db, query_db, que_ind: list, metric):
    """Logs the top hits of one query protein (reference src/query_db.py:17-59).

    ``dm`` / ``im`` are [n_query_fingerprints, k]; all entries are pooled, ordered by distance for
    metric 'l1' (stable: ties keep (row, column) order) or by decreasing value for 'ip', cut at
    ``top`` and at the first padding id (-1).  ``im`` holds 0-based index positions (= vid - 1).
    """
    flat_d = np.asarray(dm).reshape(-1)
    flat_i = np.asarray(im).reshape(-1)
    k = np.asarray(dm).shape[1] if np.asarray(dm).ndim == 2 else 1
    if metric == 'l1':       # reference: ``metric == ('l1' or 'l2')`` is true for 'l1' only
        order = np.argsort(flat_d, kind='stable')
    elif metric == 'ip':
        order = np.argsort(-flat_d, kind='stable')
    else:
        order = np.arange(flat_d.size)
    for rank, pos in enumerate(order[:top]):
        db_vid = int(flat_i[pos])
        if db_vid == -1:
            break
        q_vid = int(que_ind[pos // k])
        db_pid, db_domain = _label(fp_db, db_vid + 1)
        q_pid, q_domain = _label(query_db, q_vid)
        score = round(1 - (np.float32(flat_d[pos]) / 17000), 4)     # numpy float32, as the reference
        logging.info('Query: %s %s, Result %s: %s %s, Similarity: %s',
                     q_pid, q_domain, rank + 1, db_pid, db_domain, score)


_SELECT = """ SELECT pid, domain FROM fingerprints WHERE vid = ? """


def _label(db, vid: int):
    """(pid, domain) of a fingerprint row: the reference's SELECT (src/query_db.py:50-57), answered from the labels
    ``search_db`` fetched in bulk when there are any."""
    cache = getattr(db, '_dctd_labels', None)
    if cache is not None and vid in cache:
        return cache[vid]
    return db.cur.execute(_SELECT, (vid,)).fetchone()


def _prefetch_labels(db, vids):
    """One ``WHERE vid IN (...)`` query per 900 rows instead of one SELECT per hit (SQLite's default limit on bound
    parameters is 999).  Objects that do not accept attributes simply keep answering row by row."""
    vids = sorted({int(v) for v in vids})
    try:
        cache = getattr(db, '_dctd_labels', None)
        if cache is None:
            cache = {}
            db._dctd_labels = cache
    except AttributeError:
        return
    todo = [v for v in vids if v not in cache]
    for a in range(0, len(todo), 900):
        part = todo[a:a + 900]
        marks = ','.join('?' * len(part))
        for vid, pid, dom in db.cur.execute(f'SELECT vid, pid, domain FROM fingerprints WHERE vid IN ({marks})', part):
            cache[int(vid)] = (pid, dom)


def _open(db, database_cls):
    if not isinstance(db, str):
        return db, False
    if database_cls is None:
        try:
            from database import Database as database_cls   # the reference's src/database.py on sys.path
        except ImportError as exc:
            raise ImportError('search_db needs the reference Database class: put the reference src/ on '
                              'sys.path (with dctdomain_b200.install_faiss_shim()) or pass database_cls=') from exc
    return database_cls(db), True


def search_db(args: argparse.Namespace, query_db, fp_db, metric: str = 'l1', database_cls=None, index=None):
    """Reference src/query_db.py:62-91 with the per-protein searches batched into one GPU launch."""
    fp_path = fp_db
    query_db, close_q = _open(query_db, database_cls)
    if index is None:
        print('Loading index...\n')
        index = dindex.read_index(fp_path.replace('.db', '.index'))
    index.metric_type = dindex.METRIC_L1
    fp_db, close_f = _open(fp_db, database_cls)

    pids = query_db.cur.execute(""" SELECT pid FROM sequences """).fetchall()
    print('Querying database...\n')
    arrs, inds, spans = [], [], []
    for (pid,) in pids:
        qfps = query_db.load_fprints(pid=pid)
        spans.append((len(inds), len(inds) + len(qfps)))
        arrs += [fp[1] for fp in qfps]
        inds += [fp[0] for fp in qfps]
    if arrs:
        dm, im = index.search(np.array(arrs), args.khits)
        # labels of everything that can be printed, fetched in bulk (reference: two SELECTs per hit, :50-57)
        _prefetch_labels(fp_db, (im[im >= 0] + 1).ravel())
        _prefetch_labels(query_db, inds)
        for b, e in spans:
            if e > b:
                get_top_hits(dm[b:e], im[b:e], args.khits, fp_db, query_db, np.array(inds[b:e]), metric)
    if close_q:
        query_db.close()
    if close_f:
        fp_db.close()
