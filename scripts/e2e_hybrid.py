"""GPU box: gather kernel and copy engine at the same time (shares of the arrays), does the link take more?"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib

D, B = 1280, 512
lens = np.random.RandomState(777).randint(40, 501, size=B)
arrs = [torch.empty(int(L), D).pin_memory() for L in lens for _ in range(2)]
nbytes = sum(t.numel() * 4 for t in arrs)
dbig = torch.empty(nbytes + (1 << 20), dtype=torch.uint8, device='cuda')
L = _lib.lib()
a_src = np.array([t.data_ptr() for t in arrs], dtype=np.uint64)
a_len = np.array([t.numel() * 4 for t in arrs], dtype=np.int64)
a_off = np.concatenate([[0], np.cumsum((a_len + 255) // 256 * 256)[:-1]]).astype(np.int64)
piece = 256 << 10
s_g, s_d = torch.cuda.Stream(), torch.cuda.Stream()
for share in (0.0, 0.25, 0.5, 0.75):          # share of the arrays that go through the copy engine
    is_dma = (np.arange(len(arrs)) % 4) < round(share * 4)
    g = np.flatnonzero(~is_dma)
    d = np.flatnonzero(is_dma)
    npc = (a_len[g] + piece - 1) // piece
    total = int(npc.sum())
    owner = np.repeat(np.arange(len(g)), npc)
    first = np.cumsum(npc) - npc
    within = (np.arange(total) - first[owner]) * piece
    table = torch.empty((max(total, 1), 3), dtype=torch.int64).pin_memory()
    v = table.numpy()[:total]
    v[:, 0] = a_src[g][owner].astype(np.int64) + within
    v[:, 1] = dbig.data_ptr() + a_off[g][owner] + within
    v[:, 2] = np.minimum(a_len[g][owner] - within, piece)
    dsrc, dlen, doff = (np.ascontiguousarray(x[d]) for x in (a_src, a_len, a_off))
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        if total:
            L.dctd_h2d_gather(table.data_ptr(), total, s_g.cuda_stream)
        if len(d):
            L.dctd_h2d_rows(dsrc.ctypes.data, dlen.ctypes.data, len(d), dbig.data_ptr(), doff.ctypes.data, s_d.cuda_stream)
        s_d.synchronize()
        t1 = time.perf_counter()
        s_g.synchronize()
        t2 = time.perf_counter()
        print(f'dma share {share:.2f}: dma done {1e3 * (t1 - t0):.2f} ms, all done {1e3 * (t2 - t0):.2f} ms = {nbytes / (t2 - t0) / 1e9:.1f} GB/s')
