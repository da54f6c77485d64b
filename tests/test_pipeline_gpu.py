"""configs[0] end to end on the GPU: the 8 proteins of reference test/example.fasta (lengths; embeddings are
synthetic because ESM-2 weights are not available offline), the domain strings of test/test/example-dct.npz,
Q54YF7 (L = 1035) delivered as maxlen windows -> quantize_batch -> .npz in the reference's layout ->
IndexFlatL2 / write_index / read_index / METRIC_L1 search of all 43 vs all 43 -> dct-sim on example.pair.
Every stage is compared with the CPU oracle run on the same inputs."""
import os

import numpy as np
import pytest

import cases
import synth
from oracle import fingerprint_oracle as fo
from oracle import search_oracle as so

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


def test_configs0_pipeline(tmp_path):
    from dctdomain_b200 import dct_sim
    from dctdomain_b200 import index as dindex
    from dctdomain_b200.fingerprint import Fingerprint, quantize_batch
    fps, want_rows, sid, idx, dom = [], [], [], [0], []
    for i, (pid, (L, doms)) in enumerate(cases.EXAMPLE.items()):
        if L > 500:      # long protein: what embed_seq would stitch, handed over as windows
            case = dict(seed=300 + i, L=L, D=1280, maxlen=500, kind='esm')
            chunks = cases.stitch_chunks_for(case)
            emb = {lay: [c[lay] for c in chunks] for lay in (15, 21)}
            full = {lay: fo.stitch_chunks([c[lay] for c in chunks]) for lay in (15, 21)}
        else:
            emb = synth.layers(300 + i, L, 1280, 'esm')
            full = emb
        fps.append(Fingerprint(pid=pid, seq='A' * L, embed=emb, domains=list(doms), quants={}))
        q, kept = fo.quantize_matrix(full, list(doms), [3, 80, 3, 80])
        want_rows += [q[d] for d in kept]
    quantize_batch(fps, [3, 80, 3, 80])
    got_rows = []
    for fp in fps:
        quants = np.array([fp.quants[d] for d in fp.domains], dtype=np.int8)     # database.add_fprint, :207
        sid.append(fp.pid)
        dom += fp.domains
        idx.append(idx[-1] + len(fp.domains))
        got_rows.append(quants)
    got = np.concatenate(got_rows)
    want = np.array(want_rows).astype(np.int8)
    assert got.shape == (43, 480)
    diff = np.abs(got.astype(int) - want.astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-3

    # .npz exactly as database.save_fprints writes it (src/database.py:351-375)
    npz = str(tmp_path / 'example-dct.npz')
    np.savez(npz, sid=sid, idx=idx, dom=dom, dct=got)
    ref = np.load(os.path.join(G, 'example-dct.npz'))
    z = np.load(npz)
    assert z['dct'].dtype == np.int8 and z['dct'].shape == ref['dct'].shape and len(z['dom']) == len(ref['dom'])

    # index build / file round trip / L1 top-50 (src/database.py:241-243, src/query_db.py:75-76,87)
    index = dindex.IndexFlatL2(480)
    index.add(got)
    dindex.write_index(index, str(tmp_path / 'example.index'))
    index = dindex.read_index(str(tmp_path / 'example.index'))
    index.metric_type = dindex.METRIC_L1
    dm, im = index.search(got, 50)
    dm2, im2 = so.l1_topk(got, got, 50)
    assert np.array_equal(im, im2) and np.array_equal(dm, dm2)

    # dct-sim.py --pair example.pair (src/dct-sim.py:86-124)
    out = str(tmp_path / 'pairs.txt')
    dct_sim.main(['--dct', npz, '--pair', os.path.join(G, 'example.pair'), '--output', out])
    blocks = {s: got[idx[i]:idx[i + 1]] for i, s in enumerate(sid)}
    want_lines = ['#prot1 prot2 sim-domain sim-global']
    for line in open(os.path.join(G, 'example.pair')):
        if line[0] == '#':
            continue
        a, b = line.split()[:2]
        mx, s = so.domain_sim(blocks[a], blocks[b])
        want_lines.append(f'{a} {b} {mx} {s}')
    assert open(out).read().splitlines() == want_lines
