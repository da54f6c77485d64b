"""GPU box: where the end-to-end time of quantize_batch goes, per staging route ('gather' kernel vs one DMA per array),
and what the link gives to each route alone (no Python in between)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200.fingerprint import Fingerprint, quantize_batch

D, B = 1280, 512
lens = np.random.RandomState(777).randint(40, 501, size=B)
gen = torch.Generator().manual_seed(99)
host = [(f'p{i}', int(L), {15: torch.randn(int(L), D, generator=gen).pin_memory(),
                           21: torch.randn(int(L), D, generator=gen).pin_memory()}) for i, L in enumerate(lens)]
nbytes = sum(2 * int(L) * D * 4 for L in lens)
big = torch.empty(nbytes, dtype=torch.uint8).pin_memory()
dbig = torch.empty(nbytes, dtype=torch.uint8, device='cuda')
for _ in range(2):
    dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
t0 = time.perf_counter()
dbig.copy_(big, non_blocking=True)
torch.cuda.synchronize()
dt = time.perf_counter() - t0
print(f'one pinned copy of {nbytes / 1e9:.2f} GB: {dt * 1e3:.2f} ms = {nbytes / dt / 1e9:.1f} GB/s')

# the two routes alone: all arrays of the batch, no interpreter work between the submissions
L = _lib.lib()
arrs = [t for _, _, emb in host for t in emb.values()]
a_src = np.array([t.data_ptr() for t in arrs], dtype=np.uint64)
a_len = np.array([t.numel() * 4 for t in arrs], dtype=np.int64)
a_off = np.concatenate([[0], np.cumsum((a_len + 255) // 256 * 256)[:-1]]).astype(np.int64)
stream = torch.cuda.current_stream().cuda_stream
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    _lib.check(L.dctd_h2d_rows(a_src.ctypes.data, a_len.ctypes.data, len(a_src), dbig.data_ptr(), a_off.ctypes.data, stream), 'rows')
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    print(f'dma, {len(arrs)} arrays: submit {1e3 * (t1 - t0):.2f} ms, done {dt * 1e3:.2f} ms = {nbytes / dt / 1e9:.1f} GB/s')
for piece in (64 << 10, 256 << 10, 1 << 20):
    npc = (a_len + piece - 1) // piece
    total = int(npc.sum())
    owner = np.repeat(np.arange(len(a_len)), npc)
    first = np.cumsum(npc) - npc
    within = (np.arange(total) - first[owner]) * piece
    table = torch.empty((total, 3), dtype=torch.int64).pin_memory()
    v = table.numpy()
    v[:, 0] = a_src[owner].astype(np.int64) + within
    v[:, 1] = dbig.data_ptr() + a_off[owner] + within
    v[:, 2] = np.minimum(a_len[owner] - within, piece)
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        _lib.check(L.dctd_h2d_gather(table.data_ptr(), total, stream), 'gather')
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f'gather kernel, {total} pieces of {piece >> 10} KB: {dt * 1e3:.2f} ms = {nbytes / dt / 1e9:.1f} GB/s')

for staging in ('gather', 'dma'):
    for it in range(5):
        tm = {}
        t0 = time.perf_counter()
        fps = [Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{Ln}'], quants={}) for pid, Ln, emb in host]
        t1 = time.perf_counter()
        quantize_batch(fps, [3, 80, 3, 80], staging=staging, _timing=tm)
        t2 = time.perf_counter()
        print(f'{staging} iter {it}: objects {1e3 * (t1 - t0):.2f} ms | walk+issue {1e3 * (tm["walk_done"] - t1):.2f} | plan+launch '
              f'{1e3 * (tm["launched"] - tm["walk_done"]):.2f} | wait for copies {1e3 * (tm["copies_done"] - tm["launched"]):.2f} | kernel+D2H '
              f'{1e3 * (tm["results_on_host"] - tm["copies_done"]):.2f} | assembly {1e3 * (t2 - tm["results_on_host"]):.2f} | total {1e3 * (t2 - t0):.2f} ms'
              f' = {nbytes / (t2 - t0) / 1e9:.1f} GB/s')
