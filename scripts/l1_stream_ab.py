"""A/B of the streaming-regime search (<= 16 queries) on the GPU box, tuning build (make -C dctdomain_b200/csrc tuning;
DCTD_LIB=dctdomain_b200/libdctd_tuning.so).  Modes: 0 = fused kernel 8 consumer warps x 1 group, 6 = 4 x 2, 7 = 8 x 2,
8 = 4 x 1, 9 = the separate-launch path (sample-min, k-th, direct-load stream kernel, select)."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200 import index as dindex


def rows(n, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.clamp(torch.randn((n, 480), generator=g, device='cuda') * 27.7 + 63.6, 0, 127).round().to(torch.int8)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 6_250_000
    nqs = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 4, 8, 12, 13, 16]
    modes = [int(x) for x in sys.argv[3].split(',')] if len(sys.argv) > 3 else [9, 0, 6, 7, 8]
    L = _lib.lib()
    idx = dindex.IndexFlatL1(480)
    idx.reserve(n)
    for a in range(0, n, 1 << 20):
        idx.add(rows(min(1 << 20, n - a), 100 + a))
    res = []
    for nq in nqs:
        q = rows(nq, 7 + nq)
        out = {'nq': nq}
        ref = None
        for mode in modes:
            L.dctd_l1_set_mode(mode)
            for _ in range(3):
                r = idx.search_device(q, 50)
            torch.cuda.synchronize()
            iters = 30
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                r = idx.search_device(q, 50)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            same = True
            if ref is None:
                ref = r
            else:
                same = bool(torch.equal(ref[0], r[0]) and torch.equal(ref[1], r[1]))
            out[f'mode{mode}'] = dict(ms=round(ms, 4), GBps=round(n * 480 / ms / 1e6, 1), same=same)
        L.dctd_l1_set_mode(0)
        print(json.dumps(out), flush=True)
        res.append(out)
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump({'n': n, 'results': res}, open(f'gpurun_out/l1_stream_ab_{n}.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
