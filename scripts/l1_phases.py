"""Per-kernel times of the search paths at the sizes one rank sees (GPU box; run under
`ncu --metrics gpu__time_duration.sum --clock-control none --csv`): every case runs twice, the second pass is the
one to read.  Cases: 8192 x 1M (one GPU), 8192 x 125k with an external bound (one rank of 8 on the 1M database),
10000 x 6.25M with bound (one rank of 8 on configs[4]), 8 x 6.25M streaming, 13 x 1M reference-style call."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import index as dindex


def rows(n, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.clamp(torch.randn((n, 480), generator=g, device='cuda') * 27.7 + 63.6, 0, 127).round().to(torch.int8)


def main():
    cases = sys.argv[1].split(',') if len(sys.argv) > 1 else ['1m', 'rank8', 'stream', 'ref13']
    for case in cases:
        n, nq, g, stride = {'1m': (1_000_000, 8192, 1, 0), 'rank8': (125_000, 8192, 8, 64), 'rank2': (500_000, 8192, 2, 32),
                            'cfg4rank8': (6_250_000, 8192, 8, 64), 'stream': (6_250_000, 8, 1, 0),
                            'ref13': (1_000_000, 13, 1, 0), 'ref1': (1_000_000, 1, 1, 0)}[case]
        idx = dindex.IndexFlatL1(480)
        idx.reserve(n)
        for a in range(0, n, 1 << 20):
            idx.add(rows(min(1 << 20, n - a), 100 + a))
        q = rows(nq, 7)
        for rep in range(2):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            if g > 1:
                b = idx.bound_device(q, -(-50 // g), stride)       # the MAX over ranks is this rank's own value here
                keys = idx.search_keys_device(q, 50, bound=b)
            else:
                keys = idx.search_keys_device(q, 50)
            dindex.keys_merge(keys.view(1, nq, 50))
            e1.record()
            torch.cuda.synchronize()
            print(f'case {case} rep {rep}: {e0.elapsed_time(e1):.3f} ms', flush=True)
        del idx
        torch.cuda.empty_cache()


if __name__ == '__main__':
    main()
