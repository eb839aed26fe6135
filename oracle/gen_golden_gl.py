"""Golden vectors for the global batch scheme, exported from the reference.
TEST INFRASTRUCTURE ONLY; imported by oracle/gen_golden.py.

``Global.run`` (``_schemes.py:186-191``) decodes ONE window of ``d*n`` layers
and *counts* the logical error strings of the leftover (``:194-203``) through
``Pairs.load``.  For every seeded shot the exporter stores the reference's
error, syndrome, the correction of the canonical-order UF and of the unmodified
Macar, and the count the unmodified ``Global.get_logical_error`` returns for
the leftover.  Where three or more leftover edges meet at a node the pairing
of string ends -- and with it, occasionally, the count -- depends on the order
in which the leftover ``set`` is iterated (``_schemes.py:196-197``); the pinned
order is ascending edge id, the unmodified reference's own count on the same
leftover is stored next to it (``*_count_unmodified``).
"""
from __future__ import annotations

import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.decoders import Macar  # noqa: E402
from ref_canonical import CanonicalUF  # noqa: E402

GL_CASES = [
    # name, spec (window_height = d * n), p, shots, seed
    ('gl_sf3_phen_n4', dict(code='Surface', d=3, noise='phenomenological',
                            kwargs=dict(scheme='global batch', window_height=12)), 0.06, 40, 81),
    ('gl_sf5_phen_n3', dict(code='Surface', d=5, noise='phenomenological',
                            kwargs=dict(scheme='global batch', window_height=15)), 0.04, 24, 82),
    ('gl_sf3_cl_n3', dict(code='Surface', d=3, noise='circuit-level',
                          kwargs=dict(scheme='global batch', window_height=9)), 0.03, 30, 83),
]


def gl_case(name, spec, p, shots, seed):
    import gen_golden as G
    code = G.make_code(spec)
    scheme = code.SCHEME
    assert str(scheme) == 'global batch'
    E = len(code.EDGES)
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    uf, macar = CanonicalUF(code), Macar(code)
    random.seed(seed)
    rec = dict(err=[], uf_corr=[], macar_corr=[], uf_count=[], macar_count=[], macar_steps=[],
               uf_count_unmodified=[], macar_count_unmodified=[])
    for _ in range(shots):
        error = code.make_error(p)
        syndrome = code.get_syndrome(error)
        rec['err'].append(G.pack((eidx[e] for e in error), E))
        for key, dec in (('uf', uf), ('macar', macar)):
            dec.reset()
            step = dec.decode(set(syndrome))
            leftover = error ^ dec.correction
            scheme.reset()
            for e in sorted(leftover, key=eidx.__getitem__):       # pinned order of Global.get_logical_error
                scheme.pairs.load(e)
            a, d = code.LONG_AXIS, code.D
            count = sum(1 for u, v in scheme.pairs.as_set if abs(u[a] - v[a]) == d)
            scheme.reset()
            unmodified = scheme.get_logical_error(leftover)
            rec[f'{key}_corr'].append(G.pack((eidx[e] for e in dec.correction), E))
            rec[f'{key}_count'].append(count)
            rec[f'{key}_count_unmodified'].append(unmodified)
            if key == 'macar':
                rec['macar_steps'].append(step)
    arrays = dict(err_bits=np.array(rec['err']), uf_corr=np.array(rec['uf_corr']), macar_corr=np.array(rec['macar_corr']),
                  uf_count=np.array(rec['uf_count'], np.int32), macar_count=np.array(rec['macar_count'], np.int32),
                  macar_steps=np.array(rec['macar_steps'], np.int32),
                  uf_count_unmodified=np.array(rec['uf_count_unmodified'], np.int32),
                  macar_count_unmodified=np.array(rec['macar_count_unmodified'], np.int32))
    meta = dict(kind='global batch', spec=spec, p=p, shots=shots, seed=seed, E=E,
                n=spec['kwargs']['window_height'] // spec['d'])
    return meta, arrays


def jobs():
    return {case[0]: (lambda c=case: gl_case(*c)) for case in GL_CASES}
