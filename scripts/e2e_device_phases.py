"""Where the time of quantize_batch goes when the embeddings already live on the device (GPU box): 512 proteins,
4 contiguous domains + the global one each, CUDA tensors on Fingerprint objects; and the array API on the same data."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200.fingerprint import Fingerprint, quantize_batch, quantize_device

D, B = 1280, 512
rs = np.random.RandomState(5)
plens = rs.randint(200, 1001, size=B)
off = np.concatenate([[0], np.cumsum(plens)])
layers = [torch.randn(int(off[-1]), D, device='cuda') for _ in range(2)]
doms = []
for Lp in plens:
    cuts = np.sort(rs.choice(np.arange(30, Lp - 30, 25), size=3, replace=False))
    edges = [0] + [int(c) for c in cuts] + [int(Lp)]
    doms.append([f'{a + 1}-{b}' for a, b in zip(edges[:-1], edges[1:])] + [f'1-{Lp}'])
embeds = [{15: layers[0][off[i]:off[i + 1]], 21: layers[1][off[i]:off[i + 1]]} for i in range(B)]
nd = sum(len(d) for d in doms)
for it in range(5):
    tm = {}
    t0 = time.perf_counter()
    fps = [Fingerprint(pid=f'p{i}', seq='', embed=embeds[i], domains=list(doms[i]), quants={}) for i in range(B)]
    t1 = time.perf_counter()
    quantize_batch(fps, [3, 80, 3, 80], _timing=tm)
    t2 = time.perf_counter()
    print(f'objects iter {it}: objects {1e3 * (t1 - t0):.2f} ms | walk {1e3 * (tm["walk_done"] - t1):.2f} | parse '
          f'{1e3 * (tm["parsed"] - tm["walk_done"]):.2f} | plan+launch {1e3 * (tm["launched"] - tm["parsed"]):.2f} | kernel+D2H '
          f'{1e3 * (tm["results_on_host"] - tm["launched"]):.2f} | assembly {1e3 * (t2 - tm["results_on_host"]):.2f} | total '
          f'{1e3 * (t2 - t0):.2f} ms = {nd / (t2 - t0):.0f} fingerprints/s')
for dt in (np.int64, np.int8):
    t0 = time.perf_counter()
    for _ in range(5):
        fps = [Fingerprint(pid=f'p{i}', seq='', embed=embeds[i], domains=list(doms[i]), quants={}) for i in range(B)]
        quantize_batch(fps, [3, 80, 3, 80], quants_dtype=dt)
    print(f'objects, quants as {np.dtype(dt).name}: {5 * nd / (time.perf_counter() - t0):.0f} fingerprints/s')
for it in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(10):
        res = quantize_device(layers, off[:-1], plens, doms)
    t1 = time.perf_counter()
    host = res.fingerprints.cpu()
    t2 = time.perf_counter()
    print(f'arrays iter {it}: 10 calls issued in {1e3 * (t1 - t0):.2f} ms, done after {1e3 * (t2 - t0):.2f} ms = '
          f'{10 * nd / (t2 - t0):.0f} fingerprints/s')
