"""A/B run (GPU box): general fingerprint kernel against the warp-specialised TMA kernel with the plain longest-first
queue order and with the short items spread (the default), alternating inside one process, on five workloads:
configs[1]-shaped independent domains at D = 1280 and 640, a protein-shaped batch (4 domains + the riding global
fingerprint), short domains, and long proteins as windows.  Reports GB/s of algorithmic bytes and how many output
bytes differ from the first variant.

    python scripts/fp_ab.py [n_dom]
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200.fingerprint import execute_plan, make_plan

L = _lib.lib()


def timed(plan, srcs, out, iters=10):
    ws = torch.empty(max(plan.workspace_bytes, 256), dtype=torch.uint8, device='cuda')
    for _ in range(3):
        execute_plan(plan, srcs, out, workspace=ws)
    torch.cuda.synchronize()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(iters):
        execute_plan(plan, srcs, out, tables_resident=True, workspace=ws)
    ev[1].record()
    torch.cuda.synchronize()
    return ev[0].elapsed_time(ev[1]) / iters


def ab(name, mk, srcs, out, unique_bytes, variants, res):
    ref = None
    plans = {}
    for label, flags in variants:
        if flags not in plans:
            plans[flags] = mk(flags)
        plan = plans[flags]
        if unique_bytes is None:
            unique_bytes = plan.algorithmic_bytes
        ms = timed(plan, srcs, out)
        o = out.cpu().numpy().copy()
        if ref is None:
            ref = o
        diff = int((o != ref).sum())
        maxd = int(np.abs(o.astype(int) - ref.astype(int)).max())
        r = dict(ms=round(ms, 4), GBps=round(unique_bytes / ms / 1e6, 1), fp_per_s=round(out.shape[0] / ms * 1e3),
                 bytes_differing_from_first=diff, max_abs_diff=maxd)
        res.setdefault(name, {}).setdefault(label, []).append(r)
        print(name, label, r, flush=True)


def domains(n_dom, D, variants, res, lo=40, hi=500):
    rs = np.random.RandomState(0)
    lens = rs.randint(lo, hi + 1, size=n_dom)
    off = np.concatenate([[0], np.cumsum(lens)])
    total = int(off[-1])
    torch.manual_seed(0)
    layers = [torch.randn(total, D, device='cuda') for _ in range(2)]
    mk = lambda fl: make_plan(2, D, 3, 80, [total], [0], [1], [0] * n_dom, list(range(n_dom + 1)), off[:-1], off[1:], flags=fl)
    out = torch.empty((n_dom, 480), dtype=torch.int8, device='cuda')
    ab(f'domains_{n_dom}x{D}_L{lo}-{hi}', mk, [[layers[0]], [layers[1]]], out, None, variants, res)


def proteins(n_prot, D, variants, res):
    rs = np.random.RandomState(0)
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
    mk = lambda fl: make_plan(2, D, 3, 80, plens, list(range(n_prot)), [1] * n_prot, dom_prot, list(range(nd + 1)), sb, se, flags=fl)
    out = torch.empty((nd, 480), dtype=torch.int8, device='cuda')
    ab(f'proteins_{n_prot}x{D}_fused', mk, srcs, out, 2 * total * D * 4, variants, res)


def windows(n_prot, D, variants, res, maxlen=500, overlap=200):
    """configs[2]: long proteins delivered as maxlen windows (stride maxlen - overlap, rows in two windows averaged
    in-kernel), 3-12 contiguous domains + the global one per protein."""
    rs = np.random.RandomState(0)
    stride = maxlen - overlap
    plens = rs.randint(501, 4001, size=n_prot)
    src_rows, prot_src0, prot_nsrc = [], [], []
    for Lp in plens:
        prot_src0.append(len(src_rows))
        n = 0
        start = 0
        while True:
            rows = min(maxlen, Lp - start)
            if start > 0 and rows <= overlap:
                break
            src_rows.append(rows)
            n += 1
            if start + rows >= Lp:
                break
            start += stride
        prot_nsrc.append(n)
        # the reference drops a last window of <= 200 rows: the protein then ends with the previous window
    plens_eff = [(prot_nsrc[i] - 1) * stride + src_rows[prot_src0[i] + prot_nsrc[i] - 1] for i in range(n_prot)]
    total = int(sum(src_rows))
    torch.manual_seed(0)
    layers = [torch.randn(total, D, device='cuda') for _ in range(2)]
    off = np.concatenate([[0], np.cumsum(src_rows)])
    srcs = [[layers[l][off[i]:off[i + 1]] for i in range(len(src_rows))] for l in range(2)]
    dom_prot, sb, se = [], [], []
    for p, Lp in enumerate(plens_eff):
        k = rs.randint(3, 13)
        cuts = np.sort(rs.choice(np.arange(30, Lp - 30, 10), size=k - 1, replace=False))
        edges = [0] + [int(c) for c in cuts] + [int(Lp)]
        for a, b in zip(edges[:-1], edges[1:]):
            dom_prot.append(p); sb.append(a); se.append(b)
        dom_prot.append(p); sb.append(0); se.append(int(Lp))
    nd = len(dom_prot)
    mk = lambda fl: make_plan(2, D, 3, 80, src_rows, prot_src0, prot_nsrc, dom_prot, list(range(nd + 1)), sb, se,
                              maxlen=maxlen, overlap=overlap, flags=fl)
    out = torch.empty((nd, 480), dtype=torch.int8, device='cuda')
    ab(f'windows_{n_prot}x{D}_L501-4000', mk, srcs, out, 2 * total * D * 4, variants, res)


def main():
    n_dom = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
    variants = [('general', _lib.FP_PLAN_GENERAL_KERNEL), ('ws_longest_first', _lib.FP_PLAN_LONGEST_FIRST), ('ws_spread', 0)]
    variants = variants + variants
    res = {}
    domains(n_dom, 1280, variants, res)
    proteins(n_dom // 2, 1280, variants, res)
    domains(n_dom, 640, variants, res)
    domains(n_dom * 2, 1280, variants, res, lo=40, hi=120)
    windows(max(64, n_dom // 8), 1280, variants, res)
    for D in (1024, 480, 320):       # ProtT5 / ESM-2 t12 / t6 widths
        domains(n_dom, D, variants, res)
    proteins(n_dom // 2, 480, variants, res)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/fp_ab.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
