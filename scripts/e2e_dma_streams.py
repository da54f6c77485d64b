"""GPU box: one DMA per array spread over several streams (do the copy engines fill each other's gaps?)."""
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
for ns in (1, 2, 3, 4, 8):
    streams = [torch.cuda.Stream() for _ in range(ns)]
    parts = []
    for s in range(ns):
        sel = np.arange(s, len(arrs), ns)
        parts.append((np.ascontiguousarray(a_src[sel]), np.ascontiguousarray(a_len[sel]), np.ascontiguousarray(a_off[sel])))
    for rep in range(3):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        # interleaved submission in chunks of 16 arrays per stream
        pos = 0
        n_per = max(len(p[0]) for p in parts)
        for c0 in range(0, n_per, 16):
            for s in range(ns):
                src, ln, off = parts[s]
                c1 = min(c0 + 16, len(src))
                if c1 > c0:
                    L.dctd_h2d_rows(src[c0:c1].ctypes.data, ln[c0:c1].ctypes.data, c1 - c0, dbig.data_ptr(), off[c0:c1].ctypes.data,
                                    streams[s].cuda_stream)
        t1 = time.perf_counter()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        print(f'{ns} stream(s): submit {1e3 * (t1 - t0):.2f} ms, done {dt * 1e3:.2f} ms = {nbytes / dt / 1e9:.1f} GB/s')
