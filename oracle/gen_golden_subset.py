"""Golden vectors for subset (fixed-weight) sampling, exported from the
reference.  TEST INFRASTRUCTURE ONLY; imported by oracle/gen_golden.py.

Per case: ``Noise.subset_probabilities(p)`` (index order and both columns,
``noise/main.py:50-66``) and seeded ``Noise.force_error(weight)`` samples
(``noise/main.py:116-122``, ``noise/forcers.py:85-96``) as packed edge bits.
"""
from __future__ import annotations

import random

import numpy as np

import ref_shim

localuf = ref_shim.load()

SS_CASES = [
    # name, spec, p, weights to force, seed
    ('ss_sf3_cc', dict(code='Surface', d=3, noise='code capacity'), 0.07, [0, 1, 2, 5, 13], 91),
    ('ss_sf5_cc', dict(code='Surface', d=5, noise='code capacity'), 0.05, [1, 3, 8], 92),
    ('ss_sf3_phen', dict(code='Surface', d=3, noise='phenomenological'), 0.02, [1, 2, 4], 93),
    ('ss_rep5_cc', dict(code='Repetition', d=5, noise='code capacity'), 0.1, [1, 2, 5], 94),
    ('ss_sf3_cl_balanced', dict(code='Surface', d=3, noise='circuit-level', kwargs=dict(window_height=2)), 0.01,
     [(0, 1), (1, 0), (2, 3), (5, 2)], 95),
    ('ss_sf3_cl_standard', dict(code='Surface', d=3, noise='circuit-level',
                                kwargs=dict(window_height=2, parametrization='standard')), 0.01,
     [(1, 0, 0, 0), (0, 0, 1, 1), (2, 1, 1, 0)], 96),
]


def ss_case(name, spec, p, weights, seed):
    import gen_golden as G
    code = G.make_code(spec)
    noise = code.NOISE
    E = len(code.EDGES)
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    arrays = {}
    n_all = len(noise.ALL_WEIGHTS)
    if n_all <= 200000:
        df = noise.subset_probabilities(p)
        arrays['index'] = np.array([list(w) if isinstance(w, tuple) else [w] for w in df.index], np.int64)
        arrays['subset_prob'] = df['subset prob'].to_numpy()
        arrays['survival_prob'] = df['survival prob'].to_numpy()
    random.seed(seed)
    forced = []
    for w in weights:
        for _ in range(6):
            forced.append(G.pack((eidx[e] for e in noise.force_error(w)), E))
    arrays['forced'] = np.array(forced)
    meta = dict(kind='subset', spec=spec, p=p, weights=[list(w) if isinstance(w, tuple) else w for w in weights],
                seed=seed, E=E, n_all_weights=n_all, n_fresh=len(getattr(noise, 'FRESH_EDGES', ())))
    return meta, arrays


def jobs():
    return {case[0]: (lambda c=case: ss_case(*c)) for case in SS_CASES}
