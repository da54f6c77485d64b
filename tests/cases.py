"""Case tables shared by tests/golden/make_golden.py (reference outputs) and the parity tests."""
from __future__ import annotations

import numpy as np

import synth

# dom strings of reference test/test/example-dct.npz; lengths of reference test/example.fasta
EXAMPLE = {
    'P53875': (158, ['1-81', '82-158', '1-158']),
    'Q9VFJ2': (196, ['1-62', '63-130', '131-196', '1-196']),
    'Q6GQI8': (239, ['1-108', '109-180', '181-239', '1-239']),
    'Q9XZJ4': (244, ['1-34', '35-169', '170-244', '1-244']),
    'Q62361': (256, ['1-74', '75-208', '209-256', '1-256']),
    'Q9SWG0': (409, ['1-28', '29-149', '150-266', '267-409', '1-409']),
    'Q9VSA3': (419, ['1-34', '35-120', '121-279', '280-419', '1-419']),
    'Q54YF7': (1035, ['1-40', '41-191', '192-260', '261-345', '346-397', '398-484', '485-595',
                      '596-663', '664-772', '773-810', '811-877', '878-920', '921-1035', '1-1035']),
}


def _c(name, seed, L, D, kind, domains):
    return dict(name=name, seed=seed, L=L, D=D, kind=kind, domains=domains)


FP_CASES = []
for _i, (_pid, (_len, _doms)) in enumerate(EXAMPLE.items()):
    FP_CASES.append(_c(f'ex_{_pid}', 100 + _i, _len, 1280, 'white', _doms))
FP_CASES += [
    _c('min_L3', 1, 3, 1280, 'white', ['1-3']),
    _c('L4', 2, 4, 1280, 'white', ['1-4', '2-4']),
    _c('L22', 3, 22, 1280, 'esm', ['1-22']),
    _c('L40', 4, 40, 1280, 'white', ['1-40']),
    _c('L500', 5, 500, 1280, 'white', ['1-500', '1-250', '251-500']),
    _c('L270_white', 6, 270, 1280, 'white', ['1-270', '1-100', '101-270']),
    _c('L270_esm', 7, 270, 1280, 'esm', ['1-270', '1-100', '101-270']),
    _c('L270_offset', 8, 270, 1280, 'offset', ['1-270', '1-100', '101-270']),
    _c('D640', 9, 333, 640, 'esm', ['1-120', '121-333', '1-333']),
    _c('D80', 10, 64, 80, 'white', ['1-64', '5-30']),
    _c('D96', 11, 100, 96, 'white', ['1-100', '10-90']),
    _c('D100_scalar', 12, 77, 100, 'esm', ['1-77', '3-70']),
    _c('D2560', 13, 150, 2560, 'white', ['1-150', '20-140']),
    _c('D330_odd', 14, 90, 330, 'white', ['1-90']),
    _c('discont', 15, 400, 1280, 'esm', ['177-331,1-77', '78-176', '332-400', '1-400']),
    _c('discont3', 16, 1100, 640, 'white', ['967-1047,726-895', '1-100,200-300,1000-1100', '1-1100']),
    _c('clip_end', 17, 200, 640, 'white', ['50-300', '1-200']),
    _c('drop_seg', 18, 200, 640, 'white', ['1-100,250-300', '500-600', '1-200']),
    _c('skip_quirk', 19, 200, 640, 'white', ['300-350,310-320,1-50', '250-260,1-60,100-120']),
    _c('long_1500', 20, 1500, 640, 'esm', ['1-1500', '1-700', '701-1500']),
]

# ill-conditioned by construction (D == m: the feature-axis DCT/iDCT is an identity round trip)
ILL = {'D80'}

QDIM_CASES = [
    dict(name='q_5x44', seed=31, L=120, D=640, kind='white', domains=['1-120', '1-60'], qdim=[5, 44, 5, 44]),
    dict(name='q_mixed', seed=32, L=90, D=640, kind='esm', domains=['1-90'], qdim=[4, 64, 2, 100]),
    dict(name='q_2x16_8x128', seed=33, L=64, D=1280, kind='white', domains=['1-64', '9-40'], qdim=[2, 16, 8, 128]),
    dict(name='q_6x2', seed=34, L=50, D=96, kind='white', domains=['1-50'], qdim=[6, 2, 7, 64]),
]


def _st(name, seed, L, D, maxlen, kind, domains):
    return dict(name=name, seed=seed, L=L, D=D, maxlen=maxlen, kind=kind, domains=domains)


def _parts(seed, L, pieces):
    rs = np.random.RandomState(seed)
    doms = synth.random_partition(rs, L, pieces, min_len=22)
    # one discontinuous, unsorted domain built from the first and third parts, plus the global
    if len(doms) >= 3:
        doms = [doms[2] + ',' + doms[0]] + doms[1:2] + doms[3:]
    return doms + [f'1-{L}']


STITCH_CASES = [
    _st('st_501', 41, 501, 1280, 500, 'white', _parts(41, 501, 3)),
    _st('st_800', 42, 800, 640, 500, 'esm', _parts(42, 800, 4)),
    _st('st_801', 43, 801, 640, 500, 'white', _parts(43, 801, 5)),
    _st('st_1035', 44, 1035, 1280, 500, 'white', EXAMPLE['Q54YF7'][1]),
    _st('st_1234', 45, 1234, 640, 500, 'offset', _parts(45, 1234, 6)),
    _st('st_2500', 46, 2500, 256, 500, 'esm', _parts(46, 2500, 9)),
    _st('st_4000', 47, 4000, 256, 500, 'white', _parts(47, 4000, 12)),
    _st('st_max1000', 48, 2100, 256, 1000, 'white', _parts(48, 2100, 5)),
    _st('st_max400', 49, 1000, 256, 400, 'white', _parts(49, 1000, 4)),
]


def split_lengths(seq_len, maxlen, overlap=200):
    """Window (start, length) list of reference embedding.py:83-100 (restated in oracle/ too)."""
    if seq_len <= maxlen:
        return [(0, seq_len)]
    out = []
    for start in range(0, seq_len, maxlen - overlap):
        length = min(seq_len, start + maxlen) - start
        if length > overlap:
            out.append((start, length))
    return out


def stitch_chunks_for(case):
    """[{15: float32[len,D], 21: ...}, ...] - what the fake ESM-2 returns per window."""
    chunks = []
    for c, (_, length) in enumerate(split_lengths(case['L'], case['maxlen'])):
        chunks.append({lay: synth.embedding(case['seed'] * 1000 + c * 10 + li, length, case['D'], case['kind'])
                       for li, lay in enumerate((15, 21))})
    return chunks
