"""GPU box: the minimal drop-in - one Fingerprint.quantize(qdim) call per protein (src/make_db.py:19-33 unchanged) - on
numpy embeddings, per-call latency and where it goes."""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200.fingerprint import Fingerprint

D, B = 1280, 256
lens = np.random.RandomState(777).randint(40, 501, size=B)
rs = np.random.RandomState(1)
host = [(f'p{i}', int(L), {15: rs.standard_normal((int(L), D)).astype(np.float32),
                           21: rs.standard_normal((int(L), D)).astype(np.float32)}) for i, L in enumerate(lens)]
nbytes = sum(2 * int(L) * D * 4 for L in lens)


def run():
    for pid, L, emb in host:
        fp = Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{L}', f'1-{L // 2}', f'{L // 2 + 1}-{L}'], quants={})
        fp.quantize([3, 80, 3, 80])


run()
for rep in range(2):
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    print(f'per-protein quantize(): {dt / B * 1e3:.3f} ms per call, {B / dt:.0f} proteins/s, {3 * B / dt:.0f} fingerprints/s, {nbytes / dt / 1e9:.1f} GB/s')
pr = cProfile.Profile()
pr.enable()
run()
pr.disable()
pstats.Stats(pr).sort_stats('cumulative').print_stats(18)
