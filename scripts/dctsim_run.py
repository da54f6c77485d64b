"""GPU box: the all-pairs protein scorer on its own (n x n proteins, 2-7 fingerprints each) - the command ncu profiles
for l1_protein_kernel.    python scripts/dctsim_run.py [n_prot]"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import dct_sim

n = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
rs = np.random.RandomState(1)
sets = []
for sd in (1, 2):
    counts = rs.randint(2, 8, size=n)
    off = np.concatenate([[0], np.cumsum(counts)]).astype(np.int64)
    g = torch.Generator(device='cuda').manual_seed(sd)
    fps = torch.clamp(torch.randn((int(off[-1]), 480), generator=g, device='cuda') * 27.7 + 63.6, 0, 127).round().to(torch.int8)
    sets.append((fps, off))
for rep in range(3):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    mn, last = dct_sim.protein_scores(sets[0][0], sets[0][1], sets[1][0], sets[1][1])
    print(f'rep {rep}: {n} x {n} proteins in {(time.perf_counter() - t0) * 1e3:.1f} ms (incl. pack, D2H, widening)', flush=True)
