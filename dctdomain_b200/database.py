"""Host steps around the index that dominate wall time once the kernels are fast: vectorised
replacements for the two per-blob Python loops of the reference's ``Database`` class, taking that class (or anything
with ``cur`` and ``path``) as it is.  SQLite itself stays the reference's.

  create_index(db)        reference src/database.py:227-243   blobs -> int8 [N, 480] -> IndexFlatL2 -> <path>.index
  save_fprints(db, file)  reference src/database.py:351-375   blobs -> .npz (sid, idx, dom, dct)

The reference decodes every fingerprint with ``np.load(BytesIO(blob))`` - ~15 us per row, 15 s per million rows.  A blob
is ``np.save`` of an int8 [480] array: a fixed header followed by the 480 raw bytes, so all rows decode at once from one
joined buffer after checking that every blob has the first blob's length and header.  Anything else (mixed shapes,
foreign dtypes) falls back to the reference's row-by-row decode.
"""
from __future__ import annotations

from io import BytesIO

import numpy as np


def decode_blobs(blobs) -> np.ndarray:
    """[N] ``np.save`` blobs of equal-shape 1-D arrays -> array [N, d] of their dtype (int8 for fingerprints)."""
    n = len(blobs)
    if n == 0:
        return np.empty((0, 0), dtype=np.int8)
    first = np.load(BytesIO(blobs[0]), allow_pickle=True)
    size = len(blobs[0])
    payload = first.nbytes
    uniform = first.ndim == 1 and first.dtype != object and payload > 0 and all(len(b) == size for b in blobs)
    if uniform:
        raw = np.frombuffer(b''.join(blobs), dtype=np.uint8).reshape(n, size)
        head = size - payload
        if (raw[:, :head] == raw[0, :head]).all():                      # same header everywhere: same dtype and shape
            return np.ascontiguousarray(raw[:, head:]).view(first.dtype).reshape(n, first.shape[0])
    return np.array([np.load(BytesIO(b), allow_pickle=True) for b in blobs])


def create_index(db, faiss_module=None):
    """Reference ``Database.create_index``: every fingerprint in vid (rowid) order -> ``<db.path>.index``."""
    if faiss_module is None:
        from . import index as faiss_module
    rows = db.cur.execute(""" SELECT fingerprint FROM fingerprints """).fetchall()
    fps = np.array(decode_blobs([r[0] for r in rows]), dtype=np.int8)
    index = faiss_module.IndexFlatL2(fps.shape[1])
    index.add(fps)
    faiss_module.write_index(index, f'{db.path}.index')
    return index


def save_fprints(db, file: str):
    """Reference ``Database.save_fprints``: .npz with sid (protein ids in order of first appearance), idx (first
    fingerprint of every protein + the total), dom and dct."""
    rows = db.cur.execute(""" SELECT pid, domain, fingerprint FROM fingerprints """).fetchall()
    seqs, idxs = [], []
    seq = ''
    for i, row in enumerate(rows):
        if row[0] != seq:
            seq = row[0]
            seqs.append(seq)
            idxs.append(i)
    idxs.append(len(rows))
    doms = [r[1] for r in rows]
    fps = decode_blobs([r[2] for r in rows])
    np.savez(file, sid=seqs, idx=idxs, dom=doms, dct=fps if len(rows) else [])
