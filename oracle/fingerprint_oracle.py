"""CPU oracle for hot path 1 (DCT fingerprints).  TEST INFRASTRUCTURE - see oracle/__init__.py.

Restates, in numpy/scipy float64, what the reference computes:

  reference ``src/fingerprint.py``
    scale       :110-123   (v - min) / (max - min)
    idct_quant  :126-142   DCT-II (ortho) along the last axis of vec.T, keep ``num``
                           coefficients, length-``num`` inverse, per-row min-max, transpose
    get_doms    :145-171   "b1-e1,b2-e2" (1-indexed, inclusive) -> concatenated rows
    quantize    :174-201   per layer, per domain: idct_quant(n) -> idct_quant(m) on the
                           transpose -> reshape(n*m) -> (x*127).astype('int8')
  reference ``src/embedding.py``
    split_seq   :83-100    windows of ``maxlen`` at stride ``maxlen-overlap``, kept if len > overlap
    embed_seq   :163-187   overlap rows averaged once, (prev + cur) / 2 in float32

Two forms are provided:
  * ``*_faithful``  - same operation order as the reference (scipy.fft, per-row Python
                      ``scale`` loop).  This is the "port" timed as the CPU baseline.
  * ``*_matrix``    - the same arithmetic as two small float64 basis products; used as the
                      bulk checker.  tests/test_oracle_golden.py pins both against golden
                      outputs of the unmodified reference (tests/golden/make_golden.py).
"""
from __future__ import annotations

import numpy as np
from scipy.fft import dct, idct

OVERLAP = 200  # reference src/embedding.py:163


# --------------------------------------------------------------------------------------
# domain strings (reference src/fingerprint.py:145-171)
# --------------------------------------------------------------------------------------
def parse_domain(dom: str, n_rows: int):
    """Return ([(row_begin0, row_end_excl), ...], kept_domain_string).

    Mirrors get_doms: a segment is dropped only when its begin exceeds the number of
    rows (the ``(int(beg) or int(end)) > L`` test at fingerprint.py:166 reduces to
    ``beg > L``), an end beyond the protein is clipped by slicing, and - because the
    reference removes items from the list it is iterating - the segment that follows a
    dropped one is skipped (its rows are not used) but stays in the returned string.
    """
    parts = dom.split(',')
    segs = []
    i = 0
    while i < len(parts):
        beg, end = parts[i].split('-')
        beg, end = int(beg), int(end)
        if (beg or end) > n_rows:
            del parts[i]      # list shrinks; the reference's iterator still advances,
            i += 1            # so the element that slid into slot i is never visited
            continue
        b0 = beg - 1
        e0 = min(end, n_rows)
        if b0 < 0:            # numpy negative-index slice semantics of embed[-1:end]
            b0 = n_rows + b0
        segs.append((b0, max(e0, b0)))
        i += 1
    return segs, ','.join(parts)


def get_doms(embed: np.ndarray, dom: str):
    """float64 concatenation of the domain's rows, in listed order (fingerprint.py:160-171)."""
    segs, kept = parse_domain(dom, embed.shape[0])
    out = np.empty((0, embed.shape[1]))
    for b, e in segs:
        out = np.append(out, embed[b:e, :], axis=0)
    return out, kept


# --------------------------------------------------------------------------------------
# faithful form
# --------------------------------------------------------------------------------------
def scale(vec: np.ndarray) -> np.ndarray:
    """fingerprint.py:110-123."""
    hi = np.max(vec)
    lo = np.min(vec)
    return (vec - lo) / float(hi - lo)


def idct_quant(vec: np.ndarray, num: int) -> np.ndarray:
    """fingerprint.py:126-142."""
    coef = dct(vec.T, type=2, norm='ortho')
    back = idct(coef[:, :num], type=2, norm='ortho')
    for r in range(len(back)):
        back[r] = scale(back[r])
    return back.T


def quantize_faithful(embed: dict, domains: list, qdim: list):
    """fingerprint.py:174-201.  Returns (quants dict dom -> int array, domains list)."""
    quants: dict = {}
    with np.errstate(invalid='ignore', divide='ignore'):
        for li, layer in enumerate(embed.values()):
            n_dim, m_dim = qdim[2 * li], qdim[2 * li + 1]
            for dom in domains:
                rows, kept = get_doms(layer, dom)
                if not rows.size:
                    continue
                a = idct_quant(rows, n_dim)
                b = idct_quant(a.T, m_dim).T
                b = b.reshape(n_dim * m_dim)
                b = (b * 127).astype('int8')
                quants.setdefault(kept, []).extend(b.tolist())
    for key, val in quants.items():
        quants[key] = np.array(val)
    return quants, list(quants.keys())


# --------------------------------------------------------------------------------------
# matrix form (bulk checker)
# --------------------------------------------------------------------------------------
def _dct2_ortho(n_out: int, length: int) -> np.ndarray:
    """First n_out rows of the orthonormal DCT-II matrix of size ``length``."""
    k = np.arange(n_out, dtype=np.float64)[:, None]
    l = np.arange(length, dtype=np.float64)[None, :]
    mat = np.cos(np.pi * (2.0 * l + 1.0) * k / (2.0 * length))
    mat *= np.sqrt(2.0 / length)
    mat[0] *= np.sqrt(0.5)
    return mat


def basis(length: int, num: int) -> np.ndarray:
    """``num x length`` matrix B with idct(dct(x)[:num]) == B @ x (both ortho, type 2).

    The inverse is of length ``num`` (no zero padding), i.e. the orthonormal DCT-III of
    size ``num`` = transpose of the size-``num`` DCT-II matrix.
    """
    return _dct2_ortho(num, num).T @ _dct2_ortho(num, length)


def _minmax_rows(a: np.ndarray) -> np.ndarray:
    lo = a.min(axis=1, keepdims=True)
    hi = a.max(axis=1, keepdims=True)
    with np.errstate(invalid='ignore', divide='ignore'):
        return (a - lo) / (hi - lo)


def quant2d_matrix(rows: np.ndarray, n: int, m: int) -> np.ndarray:
    """One layer of one domain: float[L, D] -> int8[n*m] (values 0..127)."""
    x = np.asarray(rows, dtype=np.float64)
    y = basis(x.shape[0], n) @ x                 # [n, D]
    y = _minmax_rows(y.T).T                      # per feature column, over the n values
    z = y @ basis(x.shape[1], m).T               # [n, m]
    z = _minmax_rows(z)                          # per row, over the m values
    with np.errstate(invalid='ignore'):
        return (z.reshape(n * m) * 127).astype('int8')


def quantize_matrix(embed: dict, domains: list, qdim: list):
    """Same contract as quantize_faithful, via quant2d_matrix."""
    quants: dict = {}
    for li, layer in enumerate(embed.values()):
        n_dim, m_dim = qdim[2 * li], qdim[2 * li + 1]
        for dom in domains:
            rows, kept = get_doms(layer, dom)
            if not rows.size:
                continue
            quants.setdefault(kept, []).extend(quant2d_matrix(rows, n_dim, m_dim).tolist())
    for key, val in quants.items():
        quants[key] = np.array(val)
    return quants, list(quants.keys())


# --------------------------------------------------------------------------------------
# maxlen split + overlap stitch (reference src/embedding.py:83-100, 163-187)
# --------------------------------------------------------------------------------------
def split_lengths(seq_len: int, maxlen: int, overlap: int = OVERLAP):
    """[(start, length), ...] of the windows embed_seq feeds to ESM-2."""
    if seq_len <= maxlen:
        return [(0, seq_len)]
    out = []
    for start in range(0, seq_len, maxlen - overlap):
        length = min(seq_len, start + maxlen) - start
        if length > overlap:
            out.append((start, length))
    return out


def stitch_chunks(chunks: list, overlap: int = OVERLAP) -> np.ndarray:
    """Sequential float32 restatement of embedding.py:179-187 for one layer."""
    acc = np.array(chunks[0], dtype=np.float32, copy=True)
    half = np.float32(2.0)
    for cur in chunks[1:]:
        cur = np.asarray(cur, dtype=np.float32)
        acc[-overlap:] = (acc[-overlap:] + cur[:overlap]) / half
        acc = np.concatenate((acc, cur[overlap:]), axis=0)
    return acc
