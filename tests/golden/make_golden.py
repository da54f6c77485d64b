"""Generates tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py          (needs /root/reference; not run on the GPU box)

What is pinned:
  fingerprint_cases.npz  outputs of reference ``Fingerprint.quantize([3,80,3,80])``
                         (src/fingerprint.py:174-201) on seeded synthetic embeddings
                         (tests/synth.py): every (case, domain) row is the 480-value result.
  stitch_cases.npz       reference ``Embedding.embed_seq`` (src/embedding.py:153-192) driven by
                         a stub ``esm`` module and a fake model that returns seeded per-chunk
                         embeddings: sha256 of each stitched layer + the reference fingerprints
                         computed from it.
  qdim_cases.npz         non-default qdim values through the same reference code.
Inputs are NOT stored; they are regenerated from the seeds in tests/synth.py.
"""
import hashlib
import importlib.util
import json
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import synth  # noqa: E402
from cases import FP_CASES, STITCH_CASES, QDIM_CASES, stitch_chunks_for  # noqa: E402

REF = '/root/reference/src'


def load_ref(name):
    spec = importlib.util.spec_from_file_location('ref_' + name.replace('-', '_'), f'{REF}/{name}.py')
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def main():
    import warnings
    warnings.simplefilter('ignore')
    fpm = load_ref('fingerprint')

    out = {}
    meta = []
    for case in FP_CASES:
        emb = synth.layers(case['seed'], case['L'], case['D'], case['kind'])
        fp = fpm.Fingerprint(pid=case['name'], seq='A' * case['L'], embed=emb,
                             domains=list(case['domains']), quants={})
        fp.quantize([3, 80, 3, 80])
        arr = np.array([fp.quants[d] for d in fp.domains], dtype=np.int64)
        out[case['name'] + '/fp'] = arr.astype(np.int16)
        out[case['name'] + '/doms'] = np.array(fp.domains)
        meta.append(case['name'])
        print(case['name'], arr.shape)
    np.savez_compressed(os.path.join(HERE, 'fingerprint_cases.npz'), **out)

    out = {}
    for case in QDIM_CASES:
        emb = synth.layers(case['seed'], case['L'], case['D'], case['kind'])
        fp = fpm.Fingerprint(pid=case['name'], seq='A' * case['L'], embed=emb,
                             domains=list(case['domains']), quants={})
        fp.quantize(list(case['qdim']))
        out[case['name'] + '/fp'] = np.array([fp.quants[d] for d in fp.domains], dtype=np.int16)
        out[case['name'] + '/doms'] = np.array(fp.domains)
        print(case['name'], out[case['name'] + '/fp'].shape)
    np.savez_compressed(os.path.join(HERE, 'qdim_cases.npz'), **out)

    # ---- stitch: real embed_seq with a stub esm module and a fake model -----------------
    import torch
    sys.modules['esm'] = types.ModuleType('esm')
    embm = load_ref('embedding')

    class FakeModel:
        def __init__(self, chunks):
            self.chunks = chunks      # list of {layer: float32[len, D]}
            self.calls = 0

        def esm_tokenizer(self, pairs):
            (_, seq), = pairs
            return None, None, torch.zeros((1, len(seq) + 2), dtype=torch.long)

        def esm_encoder(self, tokens, repr_layers, return_contacts):
            ch = self.chunks[self.calls]
            self.calls += 1
            n = tokens.shape[1] - 2
            reps = {}
            for lay in repr_layers:
                assert ch[lay].shape[0] == n, (ch[lay].shape, n)
                pad = torch.zeros((1, n + 2, ch[lay].shape[1]), dtype=torch.float32)
                pad[0, 1:-1] = torch.from_numpy(ch[lay])
                reps[lay] = pad
            return {'representations': reps, 'contacts': torch.zeros((1, n, n))}

    out = {}
    for case in STITCH_CASES:
        chunks = stitch_chunks_for(case)
        e = embm.Embedding(pid=case['name'], seq='A' * case['L'])
        model = FakeModel(chunks)
        e.embed_seq(model, 'cpu', [15, 21], case['maxlen'])
        assert model.calls == len(chunks)
        for lay in (15, 21):
            assert e.embed[lay].shape == (case['L'], case['D']), e.embed[lay].shape
            out[f"{case['name']}/sha{lay}"] = np.array(
                hashlib.sha256(np.ascontiguousarray(e.embed[lay]).tobytes()).hexdigest())
        fp = fpm.Fingerprint(pid=case['name'], seq='A' * case['L'], embed=e.embed,
                             domains=list(case['domains']), quants={})
        fp.quantize([3, 80, 3, 80])
        out[case['name'] + '/fp'] = np.array([fp.quants[d] for d in fp.domains], dtype=np.int16)
        out[case['name'] + '/doms'] = np.array(fp.domains)
        print(case['name'], len(chunks), out[case['name'] + '/fp'].shape)
    np.savez_compressed(os.path.join(HERE, 'stitch_cases.npz'), **out)

    # ---- the reference's own shipped fixtures for the search path (copied data, not code) ----
    z = np.load('/root/reference/test/test/example-dct.npz')
    np.savez_compressed(os.path.join(HERE, 'example-dct.npz'), **{k: z[k] for k in z})
    with open('/root/reference/test/test/example-search.txt') as f:
        open(os.path.join(HERE, 'example-search.txt'), 'w').write(f.read())
    z = np.load('/root/reference/bench/G6PD/G6PD-dct.npz')
    np.savez_compressed(os.path.join(HERE, 'G6PD-dct.npz'), **{k: z[k] for k in z})
    for name in ('G6PD.pair', 'G6PD-dctsim.txt'):
        with open(f'/root/reference/bench/G6PD/{name}') as f:
            open(os.path.join(HERE, name), 'w').write(f.read())
    with open('/root/reference/test/example.pair') as f:
        open(os.path.join(HERE, 'example.pair'), 'w').write(f.read())
    json.dump({'fingerprint_cases': meta}, open(os.path.join(HERE, 'MANIFEST.json'), 'w'), indent=1)


if __name__ == '__main__':
    main()
