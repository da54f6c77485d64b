"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/dctd.h declares,
and the host-only planner works without a GPU.  No compute calls."""
import ctypes as C
import os

import numpy as np
import pytest

from dctdomain_b200 import _lib


@pytest.fixture(scope='module')
def lib():
    if not os.path.exists(_lib.LIB_PATH):
        _lib.build()
    return _lib.lib()


def test_exports_every_declared_symbol(lib):
    names = _lib.exported_symbols()
    assert len(names) >= 18
    for name in names:
        assert hasattr(lib, name), name
    assert lib.dctd_version() == 100
    assert lib.dctd_strerror(-3) == b'workspace too small'


def _plan(lib, **kw):
    from dctdomain_b200.fingerprint import make_plan
    return make_plan(**kw)


def _dump(lib, plan):
    npc, nit = C.c_int64(), C.c_int64()
    lib.dctd_fp_plan_dump(plan.handle, None, 0, None, 0, C.byref(npc), C.byref(nit))
    pieces = np.zeros((npc.value, 8), dtype=np.int32)
    items = np.zeros((nit.value, 8), dtype=np.int32)
    lib.dctd_fp_plan_dump(plan.handle, pieces.ctypes.data, npc.value, items.ctypes.data, nit.value, None, None)
    return pieces[:, :7], items


def test_planner_single_source(lib):
    plan = _plan(lib, n_layers=2, D=1280, n=3, m=80, src_rows=[300], prot_src0=[0], prot_nsrc=[1],
                 dom_prot=[0, 0], dom_seg_off=[0, 2, 3], seg_beg=[176, 0, 0], seg_end=[300, 77, 300])
    pieces, items = _dump(lib, plan)
    # domain 0 = rows 176..299 then 0..76 (listed order); domain 1 = the whole protein
    # the whole-protein domain rides on the other domain's item and only streams the rows it does not cover
    # (77..175); the protein's first row is the common pivot piece; columns: src_a,row_a,src_b,row_b,nrows,l0,g0
    assert pieces.tolist() == [[0, 77, -1, 0, 99, 77, 77], [0, 0, -1, 0, 1, 0, 0],
                               [0, 176, -1, 0, 124, 0, 176], [0, 0, -1, 0, 77, 124, 0]]
    assert len(items) == 4 and plan.algorithmic_bytes == 2 * 300 * 1280 * 4      # every row is read once
    assert items[0, 3] - items[0, 2] == 201 and items[0, 6] == 1                 # longest first; rider = domain 1
    plan2 = _plan(lib, n_layers=2, D=1280, n=3, m=80, src_rows=[300], prot_src0=[0], prot_nsrc=[1],
                  dom_prot=[0, 0], dom_seg_off=[0, 2, 3], seg_beg=[176, 0, 0], seg_end=[300, 77, 300],
                  flags=_lib.FP_PLAN_NO_FUSION)
    pieces2, items2 = _dump(lib, plan2)
    assert pieces2[:, :6].tolist() == [[0, 176, -1, 0, 124, 0], [0, 0, -1, 0, 77, 124], [0, 0, -1, 0, 300, 0]]
    assert plan2.algorithmic_bytes == 2 * (201 + 300) * 1280 * 4 and (items2[:, 6] == -1).all()


def test_planner_windows_and_split(lib):
    # L = 1234 -> windows 500, 500, 500, 334 at stride 300; rows [300c, 300c+200) averaged for c >= 1
    plan = _plan(lib, n_layers=1, D=640, n=3, m=80, src_rows=[500, 500, 500, 334], prot_src0=[0],
                 prot_nsrc=[4], dom_prot=[0], dom_seg_off=[0, 1], seg_beg=[0], seg_end=[1234])
    pieces, items = _dump(lib, plan)
    want = [[0, 0, -1, 0, 300, 0],
            [0, 300, 1, 0, 200, 300], [1, 200, -1, 0, 100, 500],
            [1, 300, 2, 0, 200, 600], [2, 200, -1, 0, 100, 800],
            [2, 300, 3, 0, 200, 900], [3, 200, -1, 0, 134, 1100]]
    assert pieces[:, :6].tolist() == want and (pieces[:, 6] == pieces[:, 5]).all()
    assert len(items) == 3 and sorted((int(a), int(b)) for a, b in items[:, 2:4]) == [(0, 412), (412, 824), (824, 1234)]
    assert plan.workspace_bytes >= 3 * 2 * 640 * 8


# word indices of the 128-byte item records of the warp-specialised kernel (csrc/fp_ws_kernel.cuh)
W = {n: i for i, n in enumerate(
    ['dom', 'layer', 'r0', 'r1', 'L', 'flags', 'split', 'nsplit', 'slab_base', 'counter', 'rider_dom', 'rider_slab',
     'rider_nsplit', 'rider_slab_base', 'rider_counter', 'Lg', 'piv_src_a', 'piv_row_a', 'piv_src_b', 'piv_row_b', 'n_runs',
     'piece_abs', 'run_src_a', 'run_row_a', 'run_src_b', 'run_row_b', 'run_rows', 'run_l0', 'run_g0'])}
F_RIDER, F_PIVOT_INLINE, F_SPLIT = 1, 2, 4


def _records(lib, plan, n_items):
    rec = np.zeros((n_items, 32), dtype=np.int32)
    assert lib.dctd_fp_plan_dump_records(plan.handle, rec.ctypes.data, n_items) == 0
    return rec


def test_planner_item_records_of_the_warp_specialised_kernel(lib):
    # the fused protein of test_planner_single_source: domain 0 = rows 176..299 + 0..76, domain 1 = the whole protein
    plan = _plan(lib, n_layers=2, D=1280, n=3, m=80, src_rows=[300], prot_src0=[0], prot_nsrc=[1],
                 dom_prot=[0, 0], dom_seg_off=[0, 2, 3], seg_beg=[176, 0, 0], seg_end=[300, 77, 300])
    _, items = _dump(lib, plan)
    rec = _records(lib, plan, len(items))
    assert (rec[:, [W['dom'], W['layer'], W['r0'], W['r1']]] == items[:, :4]).all()
    for r in rec:
        layer = r[W['layer']]
        assert (r[W['piv_src_a']], r[W['piv_row_a']], r[W['piv_src_b']]) == (0, 0, -1)       # protein row 0
        if r[W['dom']] == 0:      # the two-segment domain: carries the global fingerprint, streams two runs
            assert r[W['flags']] == F_RIDER and r[W['n_runs']] == 2 and (r[W['L']], r[W['Lg']]) == (201, 300)
            assert [r[W[k]] for k in ('run_src_a', 'run_row_a', 'run_src_b', 'run_rows', 'run_l0', 'run_g0')] == [0, 176, -1, 124, 0, 176]
            assert r[W['rider_dom']] == 1 and r[W['rider_nsplit']] == 2
            assert r[W['rider_slab_base']] == 2 * layer and r[W['rider_slab']] == 2 * layer + 1
            assert r[W['rider_counter']] == 1 + layer                  # counter 0 is the work queue
        else:                     # the global domain itself only streams the rows no other domain covers (77..175)
            assert r[W['flags']] == F_SPLIT and r[W['n_runs']] == 1 and r[W['rider_dom']] == -1
            assert [r[W[k]] for k in ('run_row_a', 'run_rows', 'run_l0', 'run_g0')] == [77, 99, 77, 77]
            assert (r[W['split']], r[W['nsplit']], r[W['slab_base']], r[W['counter']]) == (0, 2, 2 * layer, 1 + layer)
    # an unfused domain that starts at its own first row takes the pivot from its first stage
    plan = _plan(lib, n_layers=1, D=640, n=3, m=80, src_rows=[500, 500, 500, 334], prot_src0=[0],
                 prot_nsrc=[4], dom_prot=[0, 0], dom_seg_off=[0, 1, 2], seg_beg=[0, 320], seg_end=[250, 420])
    _, items = _dump(lib, plan)
    rec = _records(lib, plan, len(items))
    by_dom = {int(r[W['dom']]): r for r in rec}
    assert by_dom[0][W['flags']] == F_PIVOT_INLINE and by_dom[0][W['n_runs']] == 1
    # rows 320..419 lie in the overlap of windows 0 and 1: a dual-source run whose first row is the (dual) pivot
    d1 = by_dom[1]
    assert d1[W['flags']] == F_PIVOT_INLINE and d1[W['n_runs']] == 1
    assert [d1[W[k]] for k in ('run_src_a', 'run_row_a', 'run_src_b', 'run_row_b', 'run_rows')] == [0, 320, 1, 20, 100]
    # a domain split over several items: every item but the first needs the pivot as a separate stage
    plan = _plan(lib, n_layers=1, D=640, n=3, m=80, src_rows=[1500], prot_src0=[0], prot_nsrc=[1],
                 dom_prot=[0], dom_seg_off=[0, 1], seg_beg=[0], seg_end=[1500])
    _, items = _dump(lib, plan)
    rec = _records(lib, plan, len(items))
    assert len(rec) == 3 and (rec[:, W['nsplit']] == 3).all() and sorted(rec[:, W['split']]) == [0, 1, 2]
    for r in rec:
        assert r[W['flags']] == (F_SPLIT | (F_PIVOT_INLINE if r[W['r0']] == 0 else 0))
        assert (r[W['piv_src_a']], r[W['piv_row_a']]) == (0, 0) and r[W['run_row_a']] == r[W['r0']]


def test_planner_rejects_bad_input(lib):
    with pytest.raises(ValueError):      # domain shorter than n: the reference's reshape fails too
        _plan(lib, n_layers=1, D=640, n=3, m=80, src_rows=[10], prot_src0=[0], prot_nsrc=[1],
              dom_prot=[0], dom_seg_off=[0, 1], seg_beg=[0], seg_end=[2])
    with pytest.raises(ValueError):      # m > D
        _plan(lib, n_layers=1, D=64, n=3, m=80, src_rows=[10], prot_src0=[0], prot_nsrc=[1],
              dom_prot=[0], dom_seg_off=[0, 1], seg_beg=[0], seg_end=[10])
    with pytest.raises(_lib.DctdError):  # stride < overlap: rows would sit in three windows
        _plan(lib, n_layers=1, D=640, n=3, m=80, src_rows=[300, 300, 250], prot_src0=[0], prot_nsrc=[3],
              dom_prot=[0], dom_seg_off=[0, 1], seg_beg=[0], seg_end=[300], maxlen=300, overlap=200)


def test_parse_domain_matches_reference_quirks():
    from dctdomain_b200.fingerprint import parse_domain
    assert parse_domain('300-350,310-320,1-50', 200) == ([(0, 50)], '310-320,1-50')
    assert parse_domain('50-300', 200) == ([(49, 200)], '50-300')
    assert parse_domain('500-600', 200) == ([], '')
    # removal is BY VALUE (list.remove): an equal out-of-range segment that was passed over earlier goes first
    # (expected values: the unmodified reference get_doms, src/fingerprint.py:160-171, run on these strings)
    assert parse_domain('34-25,34-33,10-17,34-33', 33) == ([(9, 17)], '10-17,34-33')
    assert parse_domain('34-33,10-17,34-33', 33) == ([], '10-17')
    assert parse_domain('40-50,40-50,1-5,40-50,6-7', 33) == ([(0, 5)], '1-5,40-50,6-7')
    assert parse_domain('177-331,1-77', 400) == ([(176, 331), (0, 77)], '177-331,1-77')


def test_no_oracle_import_in_product():
    root = os.path.dirname(_lib.__file__)
    for dirpath, _, files in os.walk(root):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                text = open(os.path.join(dirpath, f)).read()
                assert 'import oracle' not in text and 'from oracle' not in text, f


def test_topk_workspace_sizes_host_logic(lib):
    """dctd_l1_topk_workspace_bytes runs the same host-side configuration as dctd_l1_topk: exercise the
    edge cases (empty / tiny databases, k and d limits) without a GPU."""
    for nq, n in [(1, 0), (3, 1), (13, 43), (1, 10 ** 6), (8192, 10 ** 6), (32768, 5 * 10 ** 7), (100, 5000)]:
        for k in (1, 50, 300, 992):
            assert lib.dctd_l1_topk_workspace_bytes(nq, n, 480, k) >= nq * k * 8
    assert lib.dctd_l1_topk_workspace_bytes(8, 1000, 480, 993) == 0      # k limit
    assert lib.dctd_l1_topk_workspace_bytes(8, 1000, 4096, 10) == 0      # d limit
    assert lib.dctd_l1_topk_workspace_bytes(8, 1000, 1024, 10) > 0       # wide vectors: one group per stage
    assert lib.dctd_l1_packed_bytes(33, 480) == 64 * 480 and lib.dctd_l1_packed_bytes(1, 100) == 32 * 112


def test_search_db_label_prefetch_matches_row_by_row_selects(tmp_path):
    """search_db fetches the (pid, domain) labels of all hits with WHERE vid IN (...) queries; get_top_hits then
    answers from that cache - same tuples as the reference's per-hit SELECT (src/query_db.py:50-57)."""
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    from minidb import MiniDB
    from dctdomain_b200 import query_db as qdb
    npz = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden', 'example-dct.npz'))
    db = MiniDB(str(tmp_path / 'x.db'), npz)
    n = len(npz['dom'])
    want = {vid: db.cur.execute(qdb._SELECT, (vid,)).fetchone() for vid in range(1, n + 1)}
    assert qdb._label(db, 7) == want[7]                        # no cache yet: row-by-row path
    qdb._prefetch_labels(db, list(range(1, n + 1)) * 30)       # duplicates and > 900 bound parameters' worth of input
    assert db._dctd_labels == want
    db.cur.execute('DELETE FROM fingerprints')                 # answers now come from the cache only
    assert all(qdb._label(db, vid) == want[vid] for vid in want)
    db.close()


@pytest.mark.parametrize('seed', range(6))
def test_planner_fuzz_items_cover_every_domain_row_once(lib, seed):
    """Random geometry (windows, multi-segment / overlapping domains, fusion candidates): walking the runs of every
    item record - first run from the record, later runs from the piece list clipped to the item's row range, exactly
    what the producer warp of the warp-specialised kernel does - must visit the rows of each (domain, layer) in order
    with the right protein position; for a fused protein, rider items and filler items together visit every protein
    row exactly once per layer."""
    rs = np.random.RandomState(seed)
    maxlen, overlap = 500, 200
    stride = maxlen - overlap
    src_rows, prot_src0, prot_nsrc, plens = [], [], [], []
    for p in range(12):
        prot_src0.append(len(src_rows))
        Lp = int(rs.choice([rs.randint(3, 80), rs.randint(80, 500), rs.randint(501, 2200)]))
        if Lp <= maxlen or rs.rand() < 0.3:
            src_rows.append(Lp)
        else:
            start = 0
            while True:
                rows = min(maxlen, Lp - start)
                if start > 0 and rows <= overlap:
                    break
                src_rows.append(rows)
                if start + rows >= Lp:
                    break
                start += stride
        prot_nsrc.append(len(src_rows) - prot_src0[-1])
        plens.append(src_rows[-1] if prot_nsrc[-1] == 1 else (prot_nsrc[-1] - 1) * stride + src_rows[-1])
    dom_prot, seg_off, sb, se = [], [0], [], []
    for p, Lp in enumerate(plens):
        doms = []
        if rs.rand() < 0.5 and Lp >= 12:
            cuts = sorted(rs.choice(np.arange(3, Lp - 3), size=min(rs.randint(1, 5), max(1, (Lp - 6) // 4)), replace=False))
            edges = [0] + [int(c) for c in cuts] + [Lp]
            doms = [[(a, b)] for a, b in zip(edges[:-1], edges[1:]) if b - a >= 3]
            if rs.rand() < 0.5 and len(doms) > 1:
                doms.pop(rs.randint(len(doms)))
            doms.append([(0, Lp)])
        else:
            for _ in range(rs.randint(1, 4)):
                segs = []
                for _ in range(rs.randint(1, 4)):
                    a = rs.randint(0, Lp)
                    segs.append((a, rs.randint(a + 1, Lp + 1)))
                if sum(b - a for a, b in segs) >= 3:
                    doms.append(segs)
            if not doms:
                doms.append([(0, Lp)])
        for segs in doms:
            dom_prot.append(p)
            for a, b in segs:
                sb.append(a); se.append(b)
            seg_off.append(len(sb))
    n_layers = 2
    plan = _plan(lib, n_layers=n_layers, D=1280, n=3, m=80, src_rows=src_rows, prot_src0=prot_src0, prot_nsrc=prot_nsrc,
                 dom_prot=dom_prot, dom_seg_off=seg_off, seg_beg=sb, seg_end=se, maxlen=maxlen, overlap=overlap)
    pieces, items = _dump(lib, plan)
    rec = _records(lib, plan, len(items))
    rider_targets = {int(r[W['rider_dom']]) for r in rec if r[W['flags']] & F_RIDER}     # fused global domains

    def runs_of(r):
        out = []
        for run in range(r[W['n_runs']]):
            if run == 0:
                out.append(tuple(int(r[W[k]]) for k in ('run_src_a', 'run_row_a', 'run_src_b', 'run_row_b', 'run_rows',
                                                        'run_l0', 'run_g0')))
            else:
                sa, ra, sb_, rb, nrows, l0, g0 = (int(x) for x in pieces[r[W['piece_abs']] + run])
                a, b = max(l0, int(r[W['r0']])), min(l0 + nrows, int(r[W['r1']]))
                out.append((sa, ra + a - l0, sb_, rb + a - l0, b - a, a, g0 + a - l0))
        return out

    def protein_row(p, src, row):          # protein row that (source, row) holds
        return row if prot_nsrc[p] == 1 else (src - prot_src0[p]) * stride + row

    covered = {}
    for r in rec:
        dom, layer = int(r[W['dom']]), int(r[W['layer']])
        p = dom_prot[dom]
        expect = [x for j in range(seg_off[dom], seg_off[dom + 1]) for x in range(sb[j], se[j])]   # domain rows in order
        filler = dom in rider_targets          # a fused global domain streams only the rows nobody else covers
        pos = int(r[W['r0']])
        for sa, ra, sb_, rb, nrows, l0, g0 in runs_of(r):
            assert nrows > 0 and l0 >= pos
            # what the TMA producer reads stays inside its source tensors (memcheck is not available on the GPU pool:
            # the plan's addresses are checked here instead, over random geometries)
            assert prot_src0[p] <= sa < prot_src0[p] + prot_nsrc[p] and 0 <= ra and ra + nrows <= src_rows[sa]
            if sb_ >= 0:
                assert prot_src0[p] <= sb_ < prot_src0[p] + prot_nsrc[p] and 0 <= rb and rb + nrows <= src_rows[sb_]
            if not filler:
                assert l0 == pos
            for t in range(nrows):
                prow = protein_row(p, sa, ra + t)
                assert prow == g0 + t
                if sb_ >= 0:                   # the same protein row seen through the next window
                    assert protein_row(p, sb_, rb + t) == prow and sb_ == sa + 1
                assert expect[l0 + t] == prow
                if filler or (r[W['flags']] & F_RIDER):
                    covered.setdefault((p, layer), []).append(prow)
            pos = l0 + nrows
        if not filler:
            assert pos == r[W['r1']]
        if r[W['flags']] & F_RIDER:
            assert r[W['Lg']] == plens[p] and dom_prot[int(r[W['rider_dom']])] == p
    for (p, layer), rows in covered.items():
        assert sorted(rows) == list(range(plens[p])), (p, layer)
    assert len(covered) == n_layers * len({dom_prot[d] for d in rider_targets})


def test_queue_order_spreads_short_items_and_keeps_every_item(lib):
    """Queue order of the persistent kernel (csrc/fingerprint.cu:spread_short_items): long items stay longest first, the
    short ones (< ~480 KB of input = 96 rows at D = 1280) are spread between them instead of piling up at the end of the
    launch, the last long items carry none; DCTD_FP_PLAN_LONGEST_FIRST gives the plain order; both hold the same items."""
    rs = np.random.RandomState(5)
    lens = rs.randint(20, 600, size=3000)
    off = np.concatenate([[0], np.cumsum(lens)])
    kw = dict(n_layers=1, D=1280, n=3, m=80, src_rows=[int(off[-1])], prot_src0=[0], prot_nsrc=[1],
              dom_prot=[0] * len(lens), dom_seg_off=list(range(len(lens) + 1)), seg_beg=off[:-1], seg_end=off[1:])
    _, lpt = _dump(lib, _plan(lib, flags=_lib.FP_PLAN_LONGEST_FIRST, **kw))
    _, mix = _dump(lib, _plan(lib, **kw))
    rows = lambda it: it[:, 3] - it[:, 2]
    assert (np.diff(rows(lpt)) <= 0).all()
    assert sorted(map(tuple, lpt.tolist())) == sorted(map(tuple, mix.tolist()))          # a permutation
    r = rows(mix)
    long_, short = r[r >= 96], r[r < 96]
    assert (np.diff(long_) <= 0).all() and (np.diff(short) <= 0).all() and len(short) > 100
    pos = np.nonzero(r < 96)[0]
    assert pos.max() < len(r) - 100                     # the tail is made of long items just above the balance length
    # evenly spread by rows: the share of long rows queued before the j-th short item is ~ (j + 0.5) / n_short
    cum_long = np.cumsum(np.where(r >= 96, r, 0))
    total = cum_long[pos.max()]
    frac = cum_long[pos] / total
    want = (np.arange(len(pos)) + 0.5) / len(pos)
    assert np.abs(frac - want).max() < 0.01


def test_batch_domain_parser_matches_the_line_by_line_mirror(lib):
    """dctd_parse_domains (one C call per batch) against parse_domain, the mirror of get_doms' bound handling
    (src/fingerprint.py:160-171): clipped ends, empty strings skipped, and - through the fallback - the dropped-segment
    quirk, begin 0, anything malformed."""
    from dctdomain_b200.fingerprint import parse_domain, parse_domains_batch
    rs = np.random.RandomState(3)

    def mirror(dom_lists, plen):
        dp, off, sb, se, names = [], [0], [], [], []
        for pi, doms in enumerate(dom_lists):
            for dom in doms:
                segs, kept = parse_domain(dom, plen[pi])
                if sum(e - b for b, e in segs) == 0:
                    continue
                names.append(kept); dp.append(pi)
                for b, e in segs:
                    sb.append(b); se.append(e)
                off.append(len(sb))
        return dp, off, sb, se, names

    def check(dom_lists, plen):
        got = parse_domains_batch(dom_lists, plen)
        want = mirror(dom_lists, plen)
        for g, w in zip(got[:4], want[:4]):
            assert np.asarray(g).tolist() == list(w)
        assert list(got[4]) == want[4]

    for trial in range(30):
        n_prot = rs.randint(1, 40)
        plen = rs.randint(5, 400, size=n_prot).tolist()
        dom_lists = []
        for L in plen:
            doms = []
            for _ in range(rs.randint(0, 5)):
                segs = []
                for _ in range(rs.randint(1, 4)):
                    a = rs.randint(1, L + 1)
                    b = rs.randint(1, L + 30)            # ends beyond the protein are clipped, ends below the begin are empty
                    segs.append(f'{a}-{b}')
                doms.append(','.join(segs))
            dom_lists.append(doms)
        check(dom_lists, plen)                           # regular strings only: the C path
        if trial % 3 == 0:                               # irregular strings: the fallback must give the same
            p = rs.randint(n_prot)
            L = plen[p]
            dom_lists[p] = dom_lists[p] + [f'{L + 5}-{L + 9},3-4,1-2', f'0-{min(L, 7)}', f'1-{L}']
            check(dom_lists, plen)
    check([[], ['1-5']], [10, 10])
    check([['7-3'], ['2-2']], [10, 10])                  # an empty and a one-row domain
    with pytest.raises(ValueError):
        parse_domains_batch([['1-5-7']], [10])           # the reference's split('-') fails the same way
    with pytest.raises(ValueError):
        parse_domains_batch([['a-5']], [10])


def test_hostbind_cpulist_parser_and_noop_without_topology():
    """hostbind: the sysfs cpulist syntax, and binding is a no-op that reports itself when the platform exposes no
    NUMA node for the device (the pool's boxes are single-node VMs)."""
    from dctdomain_b200 import hostbind
    assert hostbind._parse_cpulist('0-3,8,10-11\n') == {0, 1, 2, 3, 8, 10, 11}
    assert hostbind._parse_cpulist('') == set()
    assert hostbind._parse_cpulist('5') == {5}


def test_stream_turns_are_taken_in_submission_order():
    """quantize_stream's _Turn: batches queue their copies strictly in submission order, whatever order the worker
    threads reach the point in; done() of a batch that is not at turn is a no-op, wait() of a passed turn returns."""
    import random
    import threading
    import time
    from dctdomain_b200.fingerprint import _Turn
    turn, order, lock = _Turn(), [], threading.Lock()

    def worker(seq, delay):
        time.sleep(delay)
        assert turn.is_now(seq) == (turn.now == seq)
        turn.wait(seq)
        with lock:
            order.append(seq)
        turn.done(seq)
        turn.done(seq)                 # second call: the turn has moved on, nothing happens
        turn.wait(seq)                 # a passed turn does not block

    rnd = random.Random(3)
    threads = [threading.Thread(target=worker, args=(s, rnd.random() * 0.05)) for s in range(12)]
    for t in reversed(threads):
        t.start()
    for t in threads:
        t.join(timeout=10)
        assert not t.is_alive()
    assert order == list(range(12)) and turn.now == 12


def test_h2d_rows_staged_rejects_bad_arguments_without_a_device():
    """dctd_h2d_rows_staged validates before it touches CUDA: overlapping / descending destinations, bad ring geometry."""
    import ctypes as C
    from dctdomain_b200 import _lib
    L = _lib.lib()
    src = np.zeros(64, dtype=np.uint8)
    ptrs = np.array([src.ctypes.data, src.ctypes.data], dtype=np.uint64)
    lens = np.array([64, 64], dtype=np.int64)
    ring = np.zeros(4096, dtype=np.uint8)
    fake_dev = C.c_void_p(0x1000)

    def call(off, slot=1024, slots=4, threads=2, n=2):
        o = np.array(off, dtype=np.int64)
        return L.dctd_h2d_rows_staged(ptrs.ctypes.data, lens.ctypes.data, n, fake_dev, o.ctypes.data, ring.ctypes.data, slot, slots,
                                      threads, None)
    assert call([0, 32]) == _lib.ERR_ARG    # overlapping
    assert call([128, 0]) < 0               # descending
    assert call([0, 64], slot=0) < 0
    assert call([0, 64], slots=1) < 0
    assert call([0, 64], threads=0) < 0
    assert call([0, 64], n=0) == 0          # nothing to do
