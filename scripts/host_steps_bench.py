"""Before / after timing of the blob decode that dominates the reference's create_index / save_fprints at scale
(src/database.py:227-243, 351-375): per-row np.load(BytesIO(blob)) against dctdomain_b200.database.decode_blobs, on a
SQLite table with the reference's schema.  CPU only.  usage: python scripts/host_steps_bench.py [rows]"""
import json
import os
import sqlite3
import sys
import tempfile
import time
from io import BytesIO

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
import synth
from dctdomain_b200.database import decode_blobs


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    fps = synth.fingerprints(3, min(n, 100_000))
    path = os.path.join(tempfile.mkdtemp(), 'big.db')
    conn = sqlite3.connect(path)
    cur = conn.cursor()
    cur.execute('CREATE TABLE fingerprints(vid INTEGER PRIMARY KEY, domain TEXT, fingerprint BLOB, pid TEXT)')
    blobs = []
    for r in fps:
        b = BytesIO()
        np.save(b, r, allow_pickle=True)
        blobs.append(b.getvalue())
    cur.executemany('INSERT INTO fingerprints VALUES(?,?,?,?)',
                    ((i + 1, '1-100', blobs[i % len(blobs)], f'p{i // 4}') for i in range(n)))
    conn.commit()
    t0 = time.perf_counter()
    rows = cur.execute('SELECT fingerprint FROM fingerprints').fetchall()
    t_fetch = time.perf_counter() - t0
    t0 = time.perf_counter()
    ref = np.array([np.load(BytesIO(r[0]), allow_pickle=True) for r in rows], dtype=np.int8)     # the reference's loop
    t_ref = time.perf_counter() - t0
    t0 = time.perf_counter()
    ours = np.array(decode_blobs([r[0] for r in rows]), dtype=np.int8)
    t_ours = time.perf_counter() - t0
    assert np.array_equal(ref, ours)
    out = {'rows': n, 'sqlite_fetch_s': round(t_fetch, 3), 'reference_decode_s': round(t_ref, 3),
           'vectorised_decode_s': round(t_ours, 3), 'speedup': round(t_ref / t_ours, 1)}
    print(json.dumps(out))
    os.makedirs(os.path.join(ROOT, 'profiles'), exist_ok=True)
    json.dump(out, open(os.path.join(ROOT, 'profiles', 'host_steps_r2.json'), 'w'), indent=1)
    conn.close()
    os.remove(path)


if __name__ == '__main__':
    main()
