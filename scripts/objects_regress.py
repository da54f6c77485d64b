"""GPU box: does a staged (pageable) quantize_batch call slow down later quantize_batch calls on CUDA-tensor objects?"""
import cProfile
import os
import pstats
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200.fingerprint import Fingerprint, quantize_batch, quantize_stream

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
Q = [3, 80, 3, 80]


def objects(n=5):
    t0 = time.perf_counter()
    for _ in range(n):
        fps = [Fingerprint(pid=f'p{i}', seq='', embed=embeds[i], domains=list(doms[i]), quants={}) for i in range(B)]
        quantize_batch(fps, Q)
    return (time.perf_counter() - t0) / n * 1e3


objects(2)
print(f'objects path, fresh process: {objects():.2f} ms per step')
lens = np.random.RandomState(777).randint(40, 501, size=256)
host_np = [(f'p{i}', int(L), {15: np.zeros((int(L), D), np.float32), 21: np.zeros((int(L), D), np.float32)}) for i, L in enumerate(lens)]
host_pin = [(pid, L, {k: torch.from_numpy(v).pin_memory() for k, v in emb.items()}) for pid, L, emb in host_np]
mk = lambda host: [Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{L}'], quants={}) for pid, L, emb in host]
quantize_batch(mk(host_pin), Q)
print(f'after a pinned quantize_batch: {objects():.2f} ms per step')
for _ in quantize_stream((mk(host_pin) for _ in range(3)), Q, depth=3):
    pass
print(f'after a pinned quantize_stream: {objects():.2f} ms per step')
quantize_batch(mk(host_np), Q, staging='dma')
print(f'after a pageable quantize_batch, staging=dma: {objects():.2f} ms per step')
quantize_batch(mk(host_np), Q)
print(f'after a pageable quantize_batch, staged: {objects():.2f} ms per step')
lens = np.random.RandomState(777).randint(40, 501, size=512)
pin2 = [(f'p{i}', int(L), {15: torch.randn(int(L), D).pin_memory(), 21: torch.randn(int(L), D).pin_memory()}) for i, L in enumerate(lens)]
print(f'after allocating 1.43 GB of pinned arrays: {objects():.2f} ms per step')
page2 = [(pid, L, {k: np.array(v.numpy()) for k, v in emb.items()}) for pid, L, emb in pin2]
print(f'after allocating 1.43 GB of numpy copies: {objects():.2f} ms per step')
for _ in quantize_stream((mk(page2) for _ in range(7)), Q, depth=3):
    pass
print(f'after a pageable quantize_stream over 7 batches: {objects():.2f} ms per step')
del page2
print(f'after deleting the numpy copies: {objects():.2f} ms per step')
import gc
print('gc counts', gc.get_count(), 'objects tracked', len(gc.get_objects()))
gc.collect()
print(f'after gc.collect: {objects():.2f} ms per step')
pr = cProfile.Profile()
pr.enable()
objects(3)
pr.disable()
pstats.Stats(pr).sort_stats('tottime').print_stats(12)
