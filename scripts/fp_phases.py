"""Per-phase cycle breakdown of the fingerprint kernel (needs the instrumented build:
make -C dctdomain_b200/csrc timing;  DCTD_LIB=dctdomain_b200/libdctd_timing.so python scripts/fp_phases.py)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200.fingerprint import execute_plan, make_plan

NAMES = ['fetch', 'basis', 'stream', 'rider_handoff', 'meet', 'finish_total', 'rider_finish', '-', 'fin_minmax1', 'fin_fold',
         'fin_pass2a', 'fin_reduce', 'fin_pass2b', 'fin_minmax2', 'fin_out', '-']
# warp-specialised kernel (fp_ws_kernel.cuh): one sampled thread per role, cycles summed over the CTAs
WS_NAMES = ['prod_wait_free_stage', 'prod_total', 'cons_wait_full_stage', 'cons_wait_handover', 'cons_total', 'fin_idle',
            'fin_total', 'fin_stage1', 'fin_pass2a', 'fin_out', 'items', 'fin_reduce', 'fin_pass2b']


def run(name, mk, srcs, out):
    res = {}
    for label, flags in (('general', _lib.FP_PLAN_GENERAL_KERNEL), ('ws_longest_first', _lib.FP_PLAN_LONGEST_FIRST), ('ws', 0)):
        plan = mk(flags)
        ws = torch.empty(plan.workspace_bytes, dtype=torch.uint8, device='cuda')
        for _ in range(3):
            execute_plan(plan, srcs, out, workspace=ws)
        torch.cuda.synchronize()
        buf = np.zeros(16, dtype=np.int64)
        _lib.check(_lib.lib().dctd_fp_timing_read(plan.handle, ws.data_ptr(), buf.ctypes.data))
        if label == 'general':
            tot = buf[[0, 1, 2, 3, 4, 5, 6]].sum()
            rep = {n: round(float(v) / tot, 4) for n, v in zip(NAMES, buf) if n != '-'}
            rep['cycles_per_item'] = float(tot) / plan.n_items
        else:
            rep = {n: int(v) for n, v in zip(WS_NAMES, buf)}
            rep['prod_wait_frac'] = round(buf[0] / max(1, buf[1]), 4)
            rep['cons_wait_full_frac'] = round(buf[2] / max(1, buf[4]), 4)
            rep['cons_wait_handover_frac'] = round(buf[3] / max(1, buf[4]), 4)
            rep['fin_idle_frac'] = round(buf[5] / max(1, buf[6]), 4)
            items = max(1, int(buf[10]))
            rep['fin_cycles_per_item'] = {k: round(float(buf[i]) / items) for k, i in
                                          (('stage1', 7), ('pass2a', 8), ('reduce', 11), ('pass2b', 12), ('minmax_out', 9))}
        print(name, label, json.dumps(rep), flush=True)
        res[label] = rep
    return res


def main():
    D = 1280
    rs = np.random.RandomState(0)
    torch.manual_seed(0)
    # configs[1]-like independent domains
    n_dom = 4096
    lens = rs.randint(40, 501, size=n_dom)
    off = np.concatenate([[0], np.cumsum(lens)])
    layers = [torch.randn(int(off[-1]), D, device='cuda') for _ in range(2)]
    mk = lambda fl: make_plan(2, D, 3, 80, [int(off[-1])], [0], [1], [0] * n_dom, list(range(n_dom + 1)), off[:-1], off[1:], flags=fl)
    out = torch.empty((n_dom, 480), dtype=torch.int8, device='cuda')
    res = {'domains': run('domains', mk, [[layers[0]], [layers[1]]], out)}
    # protein-shaped batch (fused)
    n_prot = 1024
    plens = rs.randint(200, 1001, size=n_prot)
    poff = np.concatenate([[0], np.cumsum(plens)])
    layers = [torch.randn(int(poff[-1]), D, device='cuda') for _ in range(2)]
    dom_prot, sb, se = [], [], []
    for p, L in enumerate(plens):
        cuts = np.sort(rs.choice(np.arange(30, L - 30, 25), size=3, replace=False))
        edges = [0] + [int(c) for c in cuts] + [int(L)]
        for a, b in zip(edges[:-1], edges[1:]):
            dom_prot.append(p); sb.append(a); se.append(b)
        dom_prot.append(p); sb.append(0); se.append(int(L))
    nd = len(dom_prot)
    srcs = [[layers[l][poff[p]:poff[p + 1]] for p in range(n_prot)] for l in range(2)]
    mk = lambda fl: make_plan(2, D, 3, 80, plens, list(range(n_prot)), [1] * n_prot, dom_prot, list(range(nd + 1)), sb, se, flags=fl)
    out = torch.empty((nd, 480), dtype=torch.int8, device='cuda')
    res['proteins_fused'] = run('proteins_fused', mk, srcs, out)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump(res, open('gpurun_out/fp_phases.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
