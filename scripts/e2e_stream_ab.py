"""GPU box: quantize_batch back to back vs quantize_stream at depth 1..3 on 512 pinned proteins per batch."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200.fingerprint import Fingerprint, quantize_batch, quantize_stream

D, B = 1280, 512
lens = np.random.RandomState(777).randint(40, 501, size=B)
gen = torch.Generator().manual_seed(99)
host = [(f'p{i}', int(L), {15: torch.randn(int(L), D, generator=gen).pin_memory(),
                           21: torch.randn(int(L), D, generator=gen).pin_memory()}) for i, L in enumerate(lens)]
nbytes = sum(2 * int(L) * D * 4 for L in lens)
Q = [3, 80, 3, 80]


def mk():
    return [Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{L}'], quants={}) for pid, L, emb in host]


N = 16
for rep in range(2):
    for _ in range(3):
        quantize_batch(mk(), Q)
    t0 = time.perf_counter()
    for _ in range(N):
        quantize_batch(mk(), Q)
    dt = (time.perf_counter() - t0) / N
    print(f'back to back: {dt * 1e3:.2f} ms per batch = {nbytes / dt / 1e9:.1f} GB/s = {B / dt:.0f} fingerprints/s')
    for depth in (1, 2, 3, 4):
        for _ in quantize_stream((mk() for _ in range(4)), Q, depth=depth):
            pass
        t0 = time.perf_counter()
        for _ in quantize_stream((mk() for _ in range(N)), Q, depth=depth):
            pass
        dt = (time.perf_counter() - t0) / N
        print(f'stream depth {depth}: {dt * 1e3:.2f} ms per batch = {nbytes / dt / 1e9:.1f} GB/s = {B / dt:.0f} fingerprints/s')
