"""GPU box: protein-shaped batch through the general kernel and the warp-specialised one, repeated; prints how many
bytes differ (a race shows up as a changing, large count).   python scripts/fp_check.py [n_prot] [reps]"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200.fingerprint import execute_plan, make_plan

n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
ws_flags = int(sys.argv[3]) if len(sys.argv) > 3 else 0          # e.g. 1 = no fusion
lmax = int(sys.argv[4]) if len(sys.argv) > 4 else 1000            # longest protein (<= 512 + 30: no split domains)
D = 1280
rs = np.random.RandomState(0)
plens = rs.randint(200, lmax + 1, size=n_prot)
poff = np.concatenate([[0], np.cumsum(plens)])
torch.manual_seed(0)
layers = [torch.randn(int(poff[-1]), D, device='cuda') for _ in range(2)]
dom_prot, sb, se = [], [], []
for p, Lp in enumerate(plens):
    cuts = np.sort(rs.choice(np.arange(30, Lp - 30, 25), size=3, replace=False))
    edges = [0] + [int(c) for c in cuts] + [int(Lp)]
    for a, b in zip(edges[:-1], edges[1:]):
        dom_prot.append(p); sb.append(a); se.append(b)
    dom_prot.append(p); sb.append(0); se.append(int(Lp))
nd = len(dom_prot)
srcs = [[layers[l][poff[p]:poff[p + 1]] for p in range(n_prot)] for l in range(2)]
L = _lib.lib()
old = not hasattr(L, 'dctd_fp_plan_create_ex')
if old:        # libraries from before the per-plan flags: same entry point without flags, general kernel via the old hook
    import ctypes as C
    L.dctd_fp_plan_create_ex = lambda geo, flags, out_: L.dctd_fp_plan_create(geo, out_)
mk = lambda fl: make_plan(2, D, 3, 80, plens, list(range(n_prot)), [1] * n_prot, dom_prot, list(range(nd + 1)), sb, se, flags=fl)
out = torch.empty((nd, 480), dtype=torch.int8, device='cuda')
if old:
    L.dctd_fp_set_variant(9)
execute_plan(mk(_lib.FP_PLAN_GENERAL_KERNEL), srcs, out)
if old:
    L.dctd_fp_set_variant(0)
ref = out.cpu().numpy().astype(int)
if os.environ.get('FP_STAGES'):
    L.dctd_fp_set_variant(100 + int(os.environ['FP_STAGES']))
plan = mk(ws_flags)
split = {i for i in range(nd) if se[i] - sb[i] > 512}
for r in range(reps):
    out.fill_(77)
    execute_plan(plan, srcs, out)
    o = out.cpu().numpy().astype(int)
    d = np.abs(o - ref)
    rows = np.unique(np.nonzero(d > 1)[0])
    print(f'rep {r}: bytes differing {int((d != 0).sum())}, max {int(d.max())}, rows with |diff| > 1: {rows[:8].tolist()} '
          f'(is global: {[int(i % 5 == 4) for i in rows[:8]]}; split domain: {[int(i in split) for i in rows[:8]]}; rows of the domain: {[se[i] - sb[i] for i in rows[:8]]}; cols {np.unique(np.nonzero(d > 1)[1] // 80)[:8].tolist()})', flush=True)
