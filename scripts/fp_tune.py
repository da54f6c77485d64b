"""Tuning run (GPU box): kernel-only throughput of the fingerprint kernel per pass-1 variant."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200.fingerprint import execute_plan, make_plan


def proteins_mode():
    """Protein-shaped batch: every protein = 4 contiguous domains + the global '1-L' domain."""
    n_prot = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
    D = 1280
    rs = np.random.RandomState(0)
    plens = rs.randint(200, 1001, size=n_prot)
    poff = np.concatenate([[0], np.cumsum(plens)])
    total = int(poff[-1])
    torch.manual_seed(0)
    layers = [torch.randn(total, D, device='cuda') for _ in range(2)]
    dom_prot, seg_beg, seg_end = [], [], []
    for p, L in enumerate(plens):
        cuts = np.sort(rs.choice(np.arange(30, L - 30, 25), size=3, replace=False))
        edges = [0] + [int(c) for c in cuts] + [int(L)]
        for a, b in zip(edges[:-1], edges[1:]):
            dom_prot.append(p); seg_beg.append(a); seg_end.append(b)
        dom_prot.append(p); seg_beg.append(0); seg_end.append(int(L))
    n_dom = len(dom_prot)
    # one source per protein (views into the slab)
    srcs = [[layers[l][poff[p]:poff[p + 1]] for p in range(n_prot)] for l in range(2)]
    out = torch.empty((n_dom, 480), dtype=torch.int8, device='cuda')
    res = {}
    for fuse, var in ((1, 0), (0, 0), (1, 0), (0, 0)):
        _lib.lib().dctd_fp_set_fusion(fuse)
        _lib.lib().dctd_fp_set_variant(var)
        plan = make_plan(2, D, 3, 80, plens, list(range(n_prot)), [1] * n_prot, dom_prot, list(range(n_dom + 1)), seg_beg, seg_end)
        for _ in range(3):
            execute_plan(plan, srcs, out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        ev[0].record()
        for _ in range(10):
            execute_plan(plan, srcs, out, tables_resident=True)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / 10
        name = f'fused_v{var}' if fuse else 'unfused'
        res[name] = dict(ms=ms, fp_per_s=n_dom / ms * 1e3, read_GBps=plan.algorithmic_bytes / ms / 1e6,
                                                    unique_GBps=2 * total * D * 4 / ms / 1e6)
        print(name, res[name], flush=True)
    _lib.lib().dctd_fp_set_fusion(1)
    _lib.lib().dctd_fp_set_variant(0)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump({'n_prot': n_prot, 'n_dom': n_dom, 'results': res}, open('gpurun_out/fp_tune_proteins.json', 'w'), indent=1)


def main():
    if len(sys.argv) > 1 and sys.argv[1] == 'proteins':
        return proteins_mode()
    n_dom = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
    D = int(sys.argv[2]) if len(sys.argv) > 2 else 1280
    variants = [int(v) for v in sys.argv[3].split(',')] if len(sys.argv) > 3 else [0, 1, 2]
    rs = np.random.RandomState(0)
    lens = rs.randint(40, 501, size=n_dom)
    off = np.concatenate([[0], np.cumsum(lens)])
    total = int(off[-1])
    torch.manual_seed(0)
    layers = [torch.randn(total, D, device='cuda') for _ in range(2)]
    plan = make_plan(2, D, 3, 80, [total], [0], [1], [0] * n_dom, list(range(n_dom + 1)), off[:-1], off[1:])
    out = torch.empty((n_dom, 480), dtype=torch.int8, device='cuda')
    res = {}
    ref = None
    for v in variants:
        _lib.lib().dctd_fp_set_variant(v)
        for _ in range(3):
            execute_plan(plan, [[layers[0]], [layers[1]]], out)
        torch.cuda.synchronize()
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
        iters = 10
        ev[0].record()
        for _ in range(iters):
            execute_plan(plan, [[layers[0]], [layers[1]]], out, tables_resident=True)
        ev[1].record()
        torch.cuda.synchronize()
        ms = ev[0].elapsed_time(ev[1]) / iters
        gbs = plan.algorithmic_bytes / ms / 1e6
        res[v] = dict(ms=ms, GBps=gbs, fp_per_s=n_dom / ms * 1e3)
        o = out.cpu().numpy().copy()
        if ref is None:
            ref = o
        res[v]['same_as_first'] = bool((o == ref).all())
        print(v, res[v], flush=True)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump({'n_dom': n_dom, 'D': D, 'bytes': plan.algorithmic_bytes, 'variants': res},
              open(f'gpurun_out/fp_tune_{n_dom}_{D}.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
