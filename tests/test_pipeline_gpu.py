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


def test_device_resident_inputs_objects_and_arrays():
    """The same batch three ways - host arrays, CUDA tensors on Fingerprint objects (incl. a protein delivered as maxlen
    windows), and the array API over a padded [B, T, D] batch (`quantize_device`, what an ESM-2 forward pass leaves on
    the device) - must give identical bytes: one kernel, one set of rows.  Also: strings that need get_doms' special
    paths (a segment beyond the protein) and a string without rows take the fallback and still agree with the oracle."""
    import torch
    from dctdomain_b200.fingerprint import Fingerprint, quantize_batch, quantize_device
    rs = np.random.RandomState(77)
    lens = [57, 300, 123, 480, 33, 260, 75, 499]
    host, doms = [], []
    for i, L in enumerate(lens):
        host.append(synth.layers(700 + i, L, 1280, synth.KINDS[i % 3]))
        d = synth.random_partition(rs, L, 2 + i % 4, min_len=10)
        doms.append(d + [f'1-{L}'])
    doms[2] = doms[2] + [f'{lens[2] + 3}-{lens[2] + 9},2-11,12-30']        # dropped segment, the next one passed over
    doms[4] = doms[4] + ['9-3']                                           # no rows: skipped

    def objects(embeds):
        return [Fingerprint(pid=f'd{i}', seq='A' * L, embed=e, domains=list(d), quants={}) for i, (L, e, d) in
                enumerate(zip(lens, embeds, doms))]

    a = quantize_batch(objects(host), [3, 80, 3, 80])
    dev = [{k: torch.from_numpy(v).cuda() for k, v in e.items()} for e in host]
    b = quantize_batch(objects(dev), [3, 80, 3, 80])
    c = quantize_batch(objects(dev), [3, 80, 3, 80], quants_dtype=np.int8)
    for fa, fb, fc, e, d in zip(a, b, c, host, doms):
        assert fa.domains == fb.domains == fc.domains and list(fa.quants) == list(fb.quants)
        want, kept = fo.quantize_matrix(e, list(d), [3, 80, 3, 80])
        assert fa.domains == kept
        for k in fa.domains:
            assert fa.quants[k].dtype == np.int64 and fc.quants[k].dtype == np.int8
            assert np.array_equal(fa.quants[k], fb.quants[k]) and np.array_equal(fa.quants[k], fc.quants[k])
            assert np.abs(fa.quants[k] - want[k]).max() <= 1
    # array API: a padded batch [B, T, D] per layer with BOS / EOS rows, as the model returns it
    T = max(lens) + 2
    layers = []
    for lid in (15, 21):
        t = torch.zeros((len(lens), T, 1280), device='cuda')
        for i, L in enumerate(lens):
            t[i, 1:L + 1] = torch.from_numpy(host[i][lid]).cuda()
            t[i, 0] = 1e6                                                  # rows outside the protein must not be read
            t[i, L + 1:] = -1e6
        layers.append(t.view(len(lens) * T, 1280))
    res = quantize_device(layers, np.arange(len(lens)) * T + 1, lens, doms)
    got = res.fingerprints.cpu().numpy()
    flat = np.concatenate([np.array([fa.quants[k] for k in fa.domains]) for fa in a])
    assert got.shape == flat.shape and np.array_equal(got.astype(np.int64), flat)
    assert res.names == [k for fa in a for k in fa.domains]
    assert res.dom_prot.tolist() == [i for i, fa in enumerate(a) for _ in fa.domains]
    # a long protein as windows on the device = the same windows from the host
    case = dict(seed=811, L=1234, D=640, maxlen=500, kind='esm')
    chunks = cases.stitch_chunks_for(case)
    wins = {lay: [c[lay] for c in chunks] for lay in (15, 21)}
    dwins = {lay: [torch.from_numpy(w).cuda() for w in ws] for lay, ws in wins.items()}
    dd = ['1-400', '401-1234', '1-1234']
    x = quantize_batch([Fingerprint(pid='w', seq='A' * 1234, embed=wins, domains=list(dd), quants={})])[0]
    y = quantize_batch([Fingerprint(pid='w', seq='A' * 1234, embed=dwins, domains=list(dd), quants={})])[0]
    assert all(np.array_equal(x.quants[k], y.quants[k]) for k in dd)
