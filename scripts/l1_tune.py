"""Tuning run (GPU box): L1 top-50 throughput across query counts, threshold path vs forced heap scan."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200 import index as dindex


MODES = [int(x) for x in os.environ.get('L1_MODES', '0,1').split(',')]


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    nqs = [int(x) for x in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1, 8, 13, 16, 64, 256, 1024, 8192]
    dev = torch.device('cuda')
    g = torch.Generator(device=dev).manual_seed(1)
    db = torch.clamp(torch.randn((n, 480), generator=g, device=dev) * 27.7 + 63.6, 0, 127).round().to(torch.int8)
    idx = dindex.IndexFlatL1(480)
    idx.add(db)
    res = []
    for nq in nqs:
        rows = torch.randint(0, n, (nq,), generator=g, device=dev)
        q = torch.clamp(db[rows].to(torch.int16) + torch.randint(-3, 4, (nq, 480), generator=g, device=dev).to(torch.int16),
                        0, 127).to(torch.int8)
        out = {}
        for mode in MODES:
            _lib.lib().dctd_l1_set_mode(mode)
            for _ in range(3):
                idx.search_device(q, 50)
            torch.cuda.synchronize()
            iters = 3 if nq >= 1024 else 20
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                r = idx.search_device(q, 50)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            out[{0: 'thresh', 1: 'heap', 2: 'thresh_heap_sample', 3: 'thresh_rescan_sample', 5: 'thresh_rolled_chunk_loop'}.get(mode, f'mode{mode}')] = dict(ms=ms, pairs_per_s=nq * n / ms * 1e3, db_GBps=n * 480 / ms / 1e6)
            out['same'] = bool(out.get('same', True) and (('ref' not in out) or (torch.equal(out['ref'][0], r[0]) and torch.equal(out['ref'][1], r[1]))))
            out.setdefault('ref', r)
        out.pop('ref')
        _lib.lib().dctd_l1_set_mode(0)
        print(nq, json.dumps(out), flush=True)
        res.append(dict(nq=nq, **out))
    os.makedirs('gpurun_out', exist_ok=True)
    json.dump({'n': n, 'results': res}, open(f'gpurun_out/l1_tune_{n}.json', 'w'), indent=1)


if __name__ == '__main__':
    main()
