"""Where the end-to-end time of quantize_batch goes (GPU box): host time stamps of one call on 512 pinned proteins."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
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
print(f'one pinned copy of {nbytes / 1e9:.2f} GB: {(time.perf_counter() - t0) * 1e3:.2f} ms')
for it in range(4):
    tm = {}
    t0 = time.perf_counter()
    fps = [Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{L}'], quants={}) for pid, L, emb in host]
    t1 = time.perf_counter()
    quantize_batch(fps, [3, 80, 3, 80], _timing=tm)
    t2 = time.perf_counter()
    print(f'iter {it}: objects {1e3 * (t1 - t0):.2f} ms | walk+issue {1e3 * (tm["walk_done"] - t1):.2f} | plan+launch '
          f'{1e3 * (tm["launched"] - tm["walk_done"]):.2f} | wait for copies {1e3 * (tm["copies_done"] - tm["launched"]):.2f} | kernel+D2H '
          f'{1e3 * (tm["results_on_host"] - tm["copies_done"]):.2f} | assembly {1e3 * (t2 - tm["results_on_host"]):.2f} | total {1e3 * (t2 - t0):.2f} ms')
