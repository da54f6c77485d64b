"""Drop-in proof against the reference's OWN modules: baseline/_ref holds verbatim copies of /root/reference/src/*.py
(scripts/install_reference.py; git-ignored, shipped to the GPU box).  With ``install_faiss_shim()`` its ``import faiss``
binds to dctdomain_b200.index, and the unmodified ``Database.add_fprint / create_index / load_fprints``
(src/database.py:197-265) and ``query_db.search_db`` (src/query_db.py:62-91) run on the GPU index and reproduce the
reference's shipped ``test/test/example-search.txt``."""
import argparse
import logging
import os
import sys
import types
from multiprocessing import Lock, Value

import numpy as np
import pytest

from conftest import GOLDEN as G

pytestmark = pytest.mark.gpu
REF = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'baseline', '_ref')


@pytest.fixture()
def ref_modules():
    if not os.path.exists(os.path.join(REF, 'database.py')):
        pytest.skip('baseline/_ref not installed (scripts/install_reference.py needs /root/reference)')
    import dctdomain_b200
    saved = {k: sys.modules.get(k) for k in ('faiss', 'esm', 'database', 'query_db', 'make_db', 'embedding', 'fingerprint')}
    dctdomain_b200.install_faiss_shim()
    sys.modules['esm'] = types.ModuleType('esm')          # embedding.py imports it at module level; never called here
    for k in ('database', 'query_db', 'make_db', 'embedding', 'fingerprint'):
        sys.modules.pop(k, None)
    sys.path.insert(0, REF)
    try:
        import database
        import query_db
        import fingerprint
        assert os.path.dirname(database.__file__) == REF and os.path.dirname(query_db.__file__) == REF
        yield database, query_db, fingerprint
    finally:
        sys.path.remove(REF)
        for k, v in saved.items():
            sys.modules.pop(k, None)
            if v is not None:
                sys.modules[k] = v


def _build_db(database, fingerprint, tmp_path, name='example'):
    """A reference database made by the reference's own code from the shipped example fingerprints."""
    z = np.load(os.path.join(G, 'example-dct.npz'))
    fa = tmp_path / f'{name}.fa'
    fa.write_text(''.join(f'>{pid}\n{"A" * (10 + i)}\n' for i, pid in enumerate(z['sid'])))
    db = database.Database(str(tmp_path / f'{name}.db'), str(fa))
    lock, counter = Lock(), Value('i', 0)
    for p, pid in enumerate(z['sid']):                       # npz order = vid order of the original database
        a, b = int(z['idx'][p]), int(z['idx'][p + 1])
        doms = [str(d) for d in z['dom'][a:b]]
        fp = fingerprint.Fingerprint(pid=str(pid), seq='', domains=doms,
                                     quants={d: z['dct'][a + i].astype(np.int64) for i, d in enumerate(doms)})
        db.add_fprint(fp, lock, counter)
    return db, z


def test_reference_database_and_search_db_on_the_shim(ref_modules, tmp_path, caplog):
    database, query_db, fingerprint = ref_modules
    db, z = _build_db(database, fingerprint, tmp_path)
    db.create_index()                                        # reference code: faiss.IndexFlatL2 / add / write_index
    db.close()
    path = str(tmp_path / 'example.db')
    assert os.path.getsize(path.replace('.db', '.index')) == 4 + 4 + 8 * 3 + 1 + 4 + 8 + len(z['dct']) * 480 * 4
    with caplog.at_level(logging.INFO):
        query_db.search_db(argparse.Namespace(khits=50), path, path)      # the unmodified reference function
    lines = [r.getMessage() for r in caplog.records if r.getMessage().startswith('Query:')]
    assert lines == open(os.path.join(G, 'example-search.txt')).read().splitlines()


def test_vectorised_create_index_and_save_fprints_equal_the_reference(ref_modules, tmp_path):
    """dctdomain_b200.database.create_index / save_fprints against the reference's own methods on the same database:
    byte-identical .index, identical .npz arrays."""
    database, _, fingerprint = ref_modules
    from dctdomain_b200 import database as ddb
    db, z = _build_db(database, fingerprint, tmp_path)
    db.create_index()
    want_index = open(f'{db.path}.index', 'rb').read()
    os.remove(f'{db.path}.index')
    ddb.create_index(db)
    assert open(f'{db.path}.index', 'rb').read() == want_index
    db.save_fprints(str(tmp_path / 'ref.npz'))
    ddb.save_fprints(db, str(tmp_path / 'ours.npz'))
    a, b = np.load(tmp_path / 'ref.npz'), np.load(tmp_path / 'ours.npz')
    assert sorted(a.files) == sorted(b.files)
    for key in a.files:
        assert a[key].dtype == b[key].dtype and np.array_equal(a[key], b[key]), key
    assert np.array_equal(b['dct'], z['dct']) and list(b['sid']) == list(z['sid']) and np.array_equal(b['idx'], z['idx'])
    db.close()
