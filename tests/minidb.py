"""Tiny SQLite stand-in with the reference's schema (src/database.py:100-126), for tests only:
the SQLite layer itself stays the reference's (out of scope)."""
import sqlite3
from io import BytesIO

import numpy as np


class MiniDB:
    def __init__(self, path, npz=None):
        self.conn = sqlite3.connect(path)
        self.cur = self.conn.cursor()
        if npz is not None:
            self.cur.execute('CREATE TABLE sequences(pid TEXT PRIMARY KEY, sequence TEXT, length INTEGER, fpcount INTEGER)')
            self.cur.execute('CREATE TABLE fingerprints(vid INTEGER PRIMARY KEY, domain TEXT, fingerprint BLOB, pid TEXT)')
            vid = 0
            for p, pid in enumerate(npz['sid']):
                a, b = int(npz['idx'][p]), int(npz['idx'][p + 1])
                self.cur.execute('INSERT INTO sequences VALUES(?,?,?,?)', (str(pid), '', 0, b - a))
                for f in range(a, b):
                    vid += 1
                    blob = BytesIO()
                    np.save(blob, npz['dct'][f], allow_pickle=True)
                    self.cur.execute('INSERT INTO fingerprints VALUES(?,?,?,?)',
                                     (vid, str(npz['dom'][f]), blob.getvalue(), str(pid)))
            self.conn.commit()

    def load_fprints(self, pid=''):
        self.cur.execute('SELECT vid, fingerprint FROM fingerprints WHERE pid = ?', (pid,))
        return [(row[0], np.load(BytesIO(row[1]), allow_pickle=True)) for row in self.cur.fetchall()]

    def close(self):
        self.conn.close()
