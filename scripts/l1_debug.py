"""Small streaming-regime searches for compute-sanitizer / phase stamps (tuning build)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from dctdomain_b200 import _lib
from dctdomain_b200 import index as dindex
from dctdomain_b200.fingerprint import _workspace


def rows(n, seed):
    g = torch.Generator(device='cuda').manual_seed(seed)
    return torch.clamp(torch.randn((n, 480), generator=g, device='cuda') * 27.7 + 63.6, 0, 127).round().to(torch.int8)


def main():
    n = int(sys.argv[1])
    nqs = [int(x) for x in sys.argv[2].split(',')]
    modes = [int(x) for x in sys.argv[3].split(',')]
    L = _lib.lib()
    idx = dindex.IndexFlatL1(480)
    idx.reserve(n)
    for a in range(0, n, 1 << 20):
        idx.add(rows(min(1 << 20, n - a), 100 + a))
    for nq in nqs:
        q = rows(nq, 7 + nq)
        for mode in modes:
            if hasattr(L, 'dctd_l1_set_mode'):
                L.dctd_l1_set_mode(mode)
            for _ in range(3):
                r = idx.search_device(q, 50)
            torch.cuda.synchronize()
            msg = f'n {n} nq {nq} mode {mode}: ok'
            if mode != 9 and hasattr(L, 'dctd_l1_stream_stamps'):
                ws = _workspace(idx._dev, 1)
                st = np.zeros(8, dtype=np.uint64)
                if L.dctd_l1_stream_stamps(ws.data_ptr(), nq, n, 480, 50, st.ctypes.data) == 0:
                    d = np.diff(st.astype(np.int64)) / 1e3
                    msg += ' stamps us: ' + ' '.join(f'{n_}={v:.1f}' for n_, v in zip(
                        ['sample', 'bar1', 'kth', 'bar2', 'stream', 'bar3', 'select'], d))
            print(msg, flush=True)


if __name__ == '__main__':
    main()
