"""GPU box: timeline of quantize_stream (device events + host stamps per batch) at a given depth."""
import collections
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dctdomain_b200.fingerprint as F

D, B = 1280, 512
lens = np.random.RandomState(777).randint(40, 501, size=B)
gen = torch.Generator().manual_seed(99)
host = [(f'p{i}', int(L), {15: torch.randn(int(L), D, generator=gen).pin_memory(),
                           21: torch.randn(int(L), D, generator=gen).pin_memory()}) for i, L in enumerate(lens)]
Q = [3, 80, 3, 80]


def mk():
    return [F.Fingerprint(pid=pid, seq='', embed=emb, domains=[f'1-{L}'], quants={}) for pid, L, emb in host]


for depth in (int(a) for a in (sys.argv[1:] or ['2', '3'])):
    for _ in F.quantize_stream((mk() for _ in range(4)), Q, depth=depth):
        pass
    torch.cuda.synchronize()
    base = torch.cuda.Event(enable_timing=True)
    base.record()
    torch.cuda.synchronize()
    t_base = time.perf_counter()
    tms = [dict() for _ in range(8)]
    stamps = []

    def batches():
        for i in range(8):
            stamps.append(('make', i, time.perf_counter()))
            yield mk()

    # same as quantize_stream, with a timing dict per batch
    dev = F._device(None)
    turn = F._Turn()
    copy_stream = F._copy_streams[dev]

    def work(fps, seq):
        st = F._stream_local.by_device[dev] if getattr(F._stream_local, 'by_device', None) and dev in F._stream_local.by_device else None
        if st is None:
            F._stream_local.by_device = getattr(F._stream_local, 'by_device', {})
            st = F._stream_local.by_device[dev] = torch.cuda.Stream(device=dev, priority=-1)
        tms[seq]['start'] = time.perf_counter()
        try:
            with torch.cuda.device(dev), torch.cuda.stream(st):
                return F.quantize_batch(fps, Q, device=dev, _copy_stream=copy_stream, _turn=(turn, seq), _timing=tms[seq])
        finally:
            tms[seq]['end'] = time.perf_counter()
            turn.wait(seq)
            turn.done(seq)

    pending = collections.deque()
    got = 0
    for seq, fps in enumerate(batches()):
        pending.append(F._stream_pool.submit(work, fps, seq))
        if len(pending) >= depth:
            pending.popleft().result()
            stamps.append(('yield', got, time.perf_counter()))
            got += 1
    while pending:
        pending.popleft().result()
        stamps.append(('yield', got, time.perf_counter()))
        got += 1
    torch.cuda.synchronize()
    print(f'--- depth {depth}: total {1e3 * (time.perf_counter() - t_base):.1f} ms for 8 batches')
    for i, tm in enumerate(tms):
        h = lambda k: 1e3 * (tm[k] - t_base) if k in tm else float('nan')
        e = lambda k: base.elapsed_time(tm[k]) if k in tm else float('nan')
        print(f'batch {i}: host start {h("start"):7.1f} walk_done {h("walk_done"):7.1f} launched {h("launched"):7.1f} results {h("results_on_host"):7.1f} end {h("end"):7.1f}'
              f' | device copies {e("copies_begin_event"):7.1f} .. {e("copies_event"):7.1f} kernel done {e("kernel_event"):7.1f}')
    print(' '.join(f'{k}{i}@{1e3 * (t - t_base):.1f}' for k, i, t in stamps))
