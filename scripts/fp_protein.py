"""GPU box: the protein-shaped batch of bench.py (`protein_batch`: 4 contiguous domains + the riding global fingerprint per
protein) on its own, a few launches - the command ncu profiles for fp_ws_kernel<2, 1280, true>.

    python scripts/fp_protein.py [n_prot] [launches]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200.fingerprint import execute_plan, make_plan


def main():
    n_prot = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    launches = int(sys.argv[2]) if len(sys.argv) > 2 else 6
    D = 1280
    rs = np.random.RandomState(50)
    plens = rs.randint(200, 1001, size=n_prot)
    poff = np.concatenate([[0], np.cumsum(plens)])
    total = int(poff[-1])
    torch.manual_seed(0)
    layers = [torch.randn(total, D, device='cuda') for _ in range(2)]
    dom_prot, sb, se = [], [], []
    for p, Lp in enumerate(plens):
        cuts = np.sort(rs.choice(np.arange(30, Lp - 30, 25), size=3, replace=False))
        edges = [0] + [int(c) for c in cuts] + [int(Lp)]
        for a, b in zip(edges[:-1], edges[1:]):
            dom_prot.append(p); sb.append(a); se.append(b)
        dom_prot.append(p); sb.append(0); se.append(int(Lp))
    nd = len(dom_prot)
    srcs = [[layers[l][poff[p]:poff[p + 1]] for p in range(n_prot)] for l in range(2)]
    plan = make_plan(2, D, 3, 80, plens, list(range(n_prot)), [1] * n_prot, dom_prot, list(range(nd + 1)), sb, se)
    out = torch.empty((nd, 480), dtype=torch.int8, device='cuda')
    ws = torch.empty(max(plan.workspace_bytes, 256), dtype=torch.uint8, device='cuda')
    execute_plan(plan, srcs, out, workspace=ws)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(launches + 1)]
    ev[0].record()
    for i in range(launches):
        execute_plan(plan, srcs, out, tables_resident=True, workspace=ws)
        ev[i + 1].record()
    torch.cuda.synchronize()
    ms = [ev[i].elapsed_time(ev[i + 1]) for i in range(launches)]
    unique = 2 * total * D * 4
    print('protein batch: %d proteins, %d fingerprints, %.2f GB; ms per launch %s; best %.0f GB/s' %
          (n_prot, nd, unique / 1e9, ['%.3f' % m for m in ms], unique / min(ms) / 1e6))


if __name__ == '__main__':
    main()
