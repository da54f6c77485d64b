"""GPU box: quantize_batch on 512 proteins whose embeddings are plain numpy arrays (pageable memory: what the reference's
.cpu().numpy() leaves) vs pinned torch tensors."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200.fingerprint import Fingerprint, quantize_batch, quantize_stream

D, B = 1280, 512
lens = np.random.RandomState(777).randint(40, 501, size=B)
rs = np.random.RandomState(1)
host_np = [(f'p{i}', int(L), {15: rs.standard_normal((int(L), D)).astype(np.float32),
                              21: rs.standard_normal((int(L), D)).astype(np.float32)}) for i, L in enumerate(lens)]
host_pin = [(pid, L, {k: torch.from_numpy(v).pin_memory() for k, v in emb.items()}) for pid, L, emb in host_np]
nbytes = sum(2 * int(L) * D * 4 for L in lens)
Q = [3, 80, 3, 80]


def mk(host):
    return [Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{L}'], quants={}) for pid, L, emb in host]


ref = None
for name, host in (('pinned', host_pin), ('pageable numpy', host_np)):
    for _ in range(2):
        fps = quantize_batch(mk(host), Q)
    got = np.stack([fp.quants[fp.domains[0]] for fp in fps])
    if ref is None:
        ref = got
    assert np.array_equal(ref, got)
    t0 = time.perf_counter()
    for _ in range(6):
        quantize_batch(mk(host), Q)
    dt = (time.perf_counter() - t0) / 6
    print(f'{name}: back to back {dt * 1e3:.2f} ms per batch = {nbytes / dt / 1e9:.1f} GB/s = {B / dt:.0f} fingerprints/s')
    for _ in quantize_stream((mk(host) for _ in range(3)), Q, depth=3):
        pass
    t0 = time.perf_counter()
    for _ in quantize_stream((mk(host) for _ in range(8)), Q, depth=3):
        pass
    dt = (time.perf_counter() - t0) / 8
    print(f'{name}: stream depth 3 {dt * 1e3:.2f} ms per batch = {nbytes / dt / 1e9:.1f} GB/s = {B / dt:.0f} fingerprints/s')
