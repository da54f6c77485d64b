"""The CPU oracle against outputs of the unmodified reference (tests/golden, made by make_golden.py)
and against the reference's own shipped fixtures.  CPU only."""
import hashlib
import os

import numpy as np
import pytest

import cases
import synth
from oracle import fingerprint_oracle as fo
from oracle import search_oracle as so

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')


@pytest.fixture(scope='module')
def fp_golden():
    return np.load(os.path.join(G, 'fingerprint_cases.npz'))


@pytest.mark.parametrize('case', cases.FP_CASES, ids=[c['name'] for c in cases.FP_CASES])
def test_matrix_oracle_bit_exact_vs_reference(case, fp_golden):
    emb = synth.layers(case['seed'], case['L'], case['D'], case['kind'])
    quants, doms = fo.quantize_matrix(emb, list(case['domains']), [3, 80, 3, 80])
    assert doms == list(fp_golden[case['name'] + '/doms'])
    got = np.array([quants[d] for d in doms])
    want = fp_golden[case['name'] + '/fp']
    assert got.shape == want.shape
    # the float64 matrix form is the same arithmetic as scipy's FFT-based DCT up to ~1e-15
    # relative; a byte can differ only if x*127 sits within that of an integer
    diff = np.abs(got.astype(int) - want.astype(int))
    # D == m makes pass 2 an identity round trip: entries that are exactly 0/1 after pass 1 come
    # back as 1 -/+ 1e-16 and the truncating cast is then decided by rounding noise (cases.ILL)
    frac = 1.0 if case['name'] in cases.ILL else 1e-4
    assert diff.max() <= 1 and (diff != 0).mean() <= frac, (diff.max(), (diff != 0).sum())


@pytest.mark.parametrize('case', cases.FP_CASES[::4], ids=[c['name'] for c in cases.FP_CASES[::4]])
def test_faithful_oracle_identical_to_reference(case, fp_golden):
    emb = synth.layers(case['seed'], case['L'], case['D'], case['kind'])
    quants, doms = fo.quantize_faithful(emb, list(case['domains']), [3, 80, 3, 80])
    assert doms == list(fp_golden[case['name'] + '/doms'])
    got = np.array([quants[d] for d in doms])
    assert np.array_equal(got, fp_golden[case['name'] + '/fp'])


@pytest.mark.parametrize('case', cases.QDIM_CASES, ids=[c['name'] for c in cases.QDIM_CASES])
def test_qdim_cases(case):
    gold = np.load(os.path.join(G, 'qdim_cases.npz'))
    emb = synth.layers(case['seed'], case['L'], case['D'], case['kind'])
    for fn in (fo.quantize_matrix, fo.quantize_faithful):
        quants, doms = fn(emb, list(case['domains']), list(case['qdim']))
        got = np.array([quants[d] for d in doms])
        want = gold[case['name'] + '/fp']
        diff = np.abs(got.astype(int) - want.astype(int))
        assert got.shape == want.shape and diff.max() <= (0 if fn is fo.quantize_faithful else 1)
        assert (diff != 0).mean() <= 1e-3


@pytest.mark.parametrize('case', cases.STITCH_CASES, ids=[c['name'] for c in cases.STITCH_CASES])
def test_stitch_bit_exact_and_fingerprints(case):
    gold = np.load(os.path.join(G, 'stitch_cases.npz'))
    chunks = cases.stitch_chunks_for(case)
    assert [c[15].shape[0] for c in chunks] == [n for _, n in fo.split_lengths(case['L'], case['maxlen'])]
    emb = {}
    for lay in (15, 21):
        emb[lay] = fo.stitch_chunks([c[lay] for c in chunks])
        assert emb[lay].shape == (case['L'], case['D'])
        sha = hashlib.sha256(np.ascontiguousarray(emb[lay]).tobytes()).hexdigest()
        assert sha == str(gold[f"{case['name']}/sha{lay}"])
    quants, doms = fo.quantize_matrix(emb, list(case['domains']), [3, 80, 3, 80])
    assert doms == list(gold[case['name'] + '/doms'])
    got = np.array([quants[d] for d in doms])
    diff = np.abs(got.astype(int) - gold[case['name'] + '/fp'].astype(int))
    assert diff.max() <= 1 and (diff != 0).mean() <= 1e-4


def test_parse_domain_quirks():
    segs, kept = fo.parse_domain('300-350,310-320,1-50', 200)
    # first segment dropped, second skipped by the reference's remove-while-iterating, third used
    assert segs == [(0, 50)] and kept == '310-320,1-50'
    segs, kept = fo.parse_domain('50-300', 200)
    assert segs == [(49, 200)] and kept == '50-300'
    segs, kept = fo.parse_domain('500-600', 200)
    assert segs == [] and kept == ''


# ---------------------------------------------------------------------------------------
# search path
# ---------------------------------------------------------------------------------------
def _labels(z):
    lab = []
    for p in range(len(z['sid'])):
        for f in range(z['idx'][p], z['idx'][p + 1]):
            lab.append((str(z['sid'][p]), str(z['dom'][f])))
    return lab


def test_search_oracle_reproduces_example_search_txt():
    """reference test/test/example-search.txt: 8 proteins x top-50 against all 43 fingerprints."""
    z = np.load(os.path.join(G, 'example-dct.npz'))
    lab = _labels(z)
    lines = []
    for p in np.argsort(z['sid']):              # SELECT pid FROM sequences -> pid ascending
        rows = list(range(z['idx'][p], z['idx'][p + 1]))
        dm, im = so.l1_topk(z['dct'][rows], z['dct'], 50, threads=2)
        dm2, im2 = so.l1_topk_numpy(z['dct'][rows], z['dct'], 50)
        assert np.array_equal(dm, dm2) and np.array_equal(im, im2)
        assert (im[:, 43:] == -1).all() and (dm[:, 43:] == so.FLT_MAX).all()
        lines += so.top_hits_lines(dm, im, 50, [lab[r] for r in rows], lab)
    want = open(os.path.join(G, 'example-search.txt')).read().splitlines()
    assert lines == want


def test_c_oracle_vs_numpy_boundary_ties():
    rs = np.random.RandomState(0)
    db = rs.randint(0, 4, size=(300, 480)).astype(np.int8)     # tiny alphabet -> many ties
    db[50:60] = db[10]                                         # exact duplicates
    q = db[:17].copy()
    for k in (1, 7, 50, 299, 300, 301):
        d1, i1 = so.l1_topk(q, db, k, threads=3)
        d2, i2 = so.l1_topk_numpy(q, db, k)
        assert np.array_equal(i1, i2) and np.array_equal(d1, d2)


def test_dct_sim_oracle_reproduces_g6pd():
    """reference bench/G6PD/G6PD-dctsim.txt from G6PD-dct.npz + G6PD.pair."""
    z = np.load(os.path.join(G, 'G6PD-dct.npz'))
    fps = {str(s): z['dct'][z['idx'][i]:z['idx'][i + 1]] for i, s in enumerate(z['sid'])}
    out = ['#prot1 prot2 sim-domain sim-global']
    for line in open(os.path.join(G, 'G6PD.pair')):
        if line[0] == '#':
            continue
        a, b = line.split()[:2]
        if a in fps and b in fps:
            mx, s = so.domain_sim(fps[a], fps[b])
            out.append(f'{a} {b} {mx} {s}')
    assert out == open(os.path.join(G, 'G6PD-dctsim.txt')).read().splitlines()
