"""Golden vectors for the forward scheme (UF / Macar), exported from the
reference.  TEST INFRASTRUCTURE ONLY; imported by oracle/gen_golden.py.

``Forward.run`` (``_schemes.py:392-450``) is restated around the reference's own
methods so that every cycle can be recorded; the one loop over a Python ``set``
that feeds ``Pairs.load`` (``_schemes.py:347-348``) runs in ascending edge id.
UF is the canonical-order subclass (oracle/ref_canonical.py); Macar is the
unmodified decoder, and the unmodified ``Forward.run`` is executed on the same
seed next to it (``unmodified_m``).
"""
from __future__ import annotations

import itertools
import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.decoders import Macar  # noqa: E402
from ref_canonical import CanonicalUF  # noqa: E402

FW_CASES = [
    # name, spec, p, n, seed
    ('fw_sf3_phen', dict(code='Surface', d=3, noise='phenomenological', kwargs=dict(scheme='forward')), 0.04, 30, 71),
    ('fw_sf5_phen', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='forward')), 0.03, 12, 72),
    ('fw_sf5_cl', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='forward')), 0.012, 10, 73),
    ('fw_sf5_phen_c2b4', dict(code='Surface', d=5, noise='phenomenological',
                              kwargs=dict(scheme='forward', commit_height=2, buffer_height=4)), 0.03, 16, 74),
    ('fw_sf7_phen', dict(code='Surface', d=7, noise='phenomenological', kwargs=dict(scheme='forward')), 0.025, 4, 75),
]


def _drive(code, dec, p, n, seed, eidx):
    import gen_golden as G
    scheme = code.SCHEME
    E = len(code.EDGES)
    c = scheme._COMMIT_HEIGHT
    random.seed(seed)
    scheme.reset()
    commit_leftover = set()
    buffer_leftover = scheme._make_error_in_buffer_region(p)
    rec = dict(init=G.pack((eidx[e] for e in buffer_leftover), E), fresh=[], steps=[], commit=[], m=[], timed=[])
    cleanse = scheme.WINDOW_HEIGHT // c
    for p_inst, timed, excl in itertools.chain(itertools.repeat((p, True, False), n), [(p, False, True)],
                                               itertools.repeat((0, False, False), cleanse)):
        seen = {code.raise_edge(e, -c) for e in buffer_leftover}
        unseen = code.make_error(p_inst, exclude_future_boundary=excl)
        error = seen | unseen
        syndrome = code.get_syndrome(error) ^ scheme._get_artificial_defects(commit_leftover)
        dec.reset()
        step = dec.decode(set(syndrome))
        commit_leftover, buffer_leftover = scheme._get_leftover(error, dec.correction)
        for e in sorted(commit_leftover, key=eidx.__getitem__):      # canonical order of Forward.get_logical_error
            scheme.pairs.load(e)
        count, scheme.pairs = scheme._LOGICAL_COUNTER.count(scheme.pairs)
        rec['fresh'].append(G.pack((eidx[e] for e in unseen), E))
        rec['steps'].append(step if step is not None else (dec.round_count, 0))
        rec['commit'].append(G.pack((eidx[e] for e in commit_leftover), E))
        rec['m'].append(count)
        rec['timed'].append(int(timed))
    return rec


def fw_case(name, spec, p, n, seed):
    import gen_golden as G
    code = G.make_code(spec)
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    arrays, ms = {}, {}
    for key, dec in (('uf', CanonicalUF(code)), ('macar', Macar(code))):
        r = _drive(code, dec, p, n, seed, eidx)
        arrays[f'{key}_init'] = r['init']
        arrays[f'{key}_fresh'] = np.array(r['fresh'])
        arrays[f'{key}_steps'] = np.array(r['steps'], np.int32)
        arrays[f'{key}_commit'] = np.array(r['commit'])
        arrays[f'{key}_m'] = np.array(r['m'], np.int32)
        arrays['timed'] = np.array(r['timed'], np.uint8)
        ms[key] = int(sum(r['m']))
    assert np.array_equal(arrays['uf_fresh'], arrays['macar_fresh'])
    random.seed(seed)
    m_unmod, slenderness = code.SCHEME.run(Macar(code), p, n)
    meta = dict(kind='forward', spec=spec, p=p, n=n, seed=seed, E=len(code.EDGES), cycles=len(arrays['timed']),
                m=ms, unmodified_macar_m=int(m_unmod), slenderness=slenderness,
                unmodified_macar_steps=[list(map(int, s)) for s in code.SCHEME.step_counts])
    return meta, arrays


def jobs():
    return {case[0]: (lambda c=case: fw_case(*c)) for case in FW_CASES}
