"""Blob decode of dctdomain_b200.database (no GPU needed): identical to the reference's per-row np.load."""
from io import BytesIO

import numpy as np

import synth


def _blob(a):
    b = BytesIO()
    np.save(b, a, allow_pickle=True)
    return b.getvalue()


def test_decode_blobs_equals_row_by_row():
    from dctdomain_b200.database import decode_blobs
    fps = synth.fingerprints(1, 257)
    blobs = [_blob(r) for r in fps]
    out = decode_blobs(blobs)
    assert out.dtype == np.int8 and np.array_equal(out, fps)
    assert decode_blobs([]).shape == (0, 0)
    # int64 rows (what quantize() returns before the cast) and ragged input fall back / decode correctly as well
    out64 = decode_blobs([_blob(r.astype(np.int64)) for r in fps[:5]])
    assert out64.dtype == np.int64 and np.array_equal(out64, fps[:5])
    mixed = [_blob(fps[0]), _blob(fps[1].astype(np.int16))]
    assert [list(x) for x in decode_blobs(mixed)] == [list(fps[0]), list(fps[1])]
