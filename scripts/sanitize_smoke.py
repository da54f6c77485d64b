"""GPU box: one small invocation of every kernel family, for compute-sanitizer:

    compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck python scripts/sanitize_smoke.py

fingerprint kernels (warp-specialised with and without riders, windows, a split domain; the general kernel), the search
paths (heap scan, threshold scan + select, fused streaming kernel, sharded keys + merge), the protein-pair scorer and
the staging helpers.  Results are checked against the oracle so that a sanitizer run is also a parity run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, 'tests')):
    sys.path.insert(0, p)
import synth  # noqa: E402
from dctdomain_b200 import _lib, dct_sim  # noqa: E402
from dctdomain_b200 import index as dindex  # noqa: E402
from dctdomain_b200.fingerprint import Fingerprint, quantize_batch, quantize_device  # noqa: E402
from oracle import fingerprint_oracle as fo  # noqa: E402
from oracle import search_oracle as so  # noqa: E402

small = os.environ.get('DCTD_SANITIZE_SMALL') == '1'        # racecheck: an order of magnitude slower


def fingerprints():
    rs = np.random.RandomState(1)
    fps, want = [], []
    for i, L in enumerate([40, 333, 90] if small else [40, 333, 90, 700, 1234]):
        for D in ((1280,) if small else (1280, 640)):
            if L > 1000:
                chunks = [synth.layers(50 + i, r, D, 'esm') for r in (500, 500, 500, 334)]
                emb = {lay: [c[lay] for c in chunks] for lay in (15, 21)}
                full = {lay: fo.stitch_chunks(emb[lay]) for lay in (15, 21)}
            else:
                emb = synth.layers(50 + i, L, D, synth.KINDS[i % 3])
                full = emb
            doms = synth.random_partition(rs, L, 2 + i % 3, min_len=12) + [f'1-{L}']
            fps.append(Fingerprint(pid=f's{i}_{D}', seq='A' * L, embed=emb, domains=list(doms), quants={}))
            want.append(fo.quantize_matrix(full, list(doms), [3, 80, 3, 80])[0])
    groups = {}
    for fp, w in zip(fps, want):
        groups.setdefault(next(iter(fp.embed.values()))[0].shape[-1] if isinstance(next(iter(fp.embed.values())), list)
                          else next(iter(fp.embed.values())).shape[-1], []).append((fp, w))
    for D, items in groups.items():
        for flags in (0, _lib.FP_PLAN_NO_FUSION, _lib.FP_PLAN_GENERAL_KERNEL):
            batch = [Fingerprint(pid=f.pid, seq=f.seq, embed=f.embed, domains=list(f.domains), quants={}) for f, _ in items]
            quantize_batch(batch, [3, 80, 3, 80], plan_flags=flags)
            for fp, (_, w) in zip(batch, items):
                for k in fp.domains:
                    assert np.abs(fp.quants[k] - w[k]).max() <= 1, (fp.pid, k, flags)
    print('fingerprint kernels ok')


def search():
    n = 20_000 if small else 70_000
    db = synth.fingerprints(5, n)
    idx = dindex.IndexFlatL1(480)
    idx.add(db)
    for nq in ((3, 40) if small else (3, 13, 40, 200)):
        q = np.concatenate([db[:nq // 2 + 1], synth.fingerprints(6, nq)])[:nq]
        dm, im = idx.search(q, 50)
        dm2, im2 = so.l1_topk(q, db, 50, threads=4)
        assert np.array_equal(im, im2) and np.array_equal(dm, dm2), nq
    # sharded pieces: bound, keys, merge
    q = torch.from_numpy(db[:64]).cuda()
    halves = []
    for part in (0, 1):
        ix = dindex.IndexFlatL1(480)
        ix.add(db[part * (n // 2):(part + 1) * (n // 2)])
        halves.append(ix.search_keys_device(q, 50, id_base=part * (n // 2)))
    dm, im = dindex.keys_merge(torch.stack(halves))
    dm2, im2 = so.l1_topk(db[:64], db, 50, threads=4)
    assert np.array_equal(im.cpu().numpy(), im2) and np.array_equal(dm.cpu().numpy(), dm2)
    print('search kernels ok')


def proteins():
    rs = np.random.RandomState(3)
    cq, cd = rs.randint(1, 12, size=60), rs.randint(1, 12, size=90)
    qoff, doff = np.concatenate([[0], np.cumsum(cq)]), np.concatenate([[0], np.cumsum(cd)])
    qf, df = synth.fingerprints(7, int(qoff[-1])), synth.fingerprints(8, int(doff[-1]))
    mn, last = dct_sim.protein_scores(qf, qoff, df, doff)
    dist = np.abs(qf[:, None, :].astype(np.int32) - df[None, :, :].astype(np.int32)).sum(axis=2)
    for a in range(60):
        for b in range(90):
            blk = dist[qoff[a]:qoff[a + 1], doff[b]:doff[b + 1]]
            assert mn[a, b] == blk.min() and last[a, b] == blk[-1, -1]
    print('protein scorer ok')


def device_api():
    lens = [50, 120, 77]
    T = max(lens) + 2
    layers = [torch.randn(len(lens) * T, 1280, device='cuda') for _ in range(2)]
    res = quantize_device(layers, np.arange(len(lens)) * T + 1, lens, [[f'1-{L // 2}', f'{L // 2 + 1}-{L}', f'1-{L}'] for L in lens])
    assert res.fingerprints.shape == (9, 480)
    torch.cuda.synchronize()
    print('device api ok')


if __name__ == '__main__':
    fingerprints()
    search()
    proteins()
    device_api()
