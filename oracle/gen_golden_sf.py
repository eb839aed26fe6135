"""Golden vectors for Snowflake under the frugal scheme, exported from the
UNMODIFIED reference decoder.  TEST INFRASTRUCTURE ONLY; imported by
oracle/gen_golden.py.

The reference ``Snowflake`` is deterministic.  ``Frugal._raise_window``
(``_schemes.py:532-543``) iterates a Python ``set`` of edges when it hands the
error edges of the floor sheet to ``Pairs.load``; string pairing at a node of
degree > 2 depends on that order (and through it, rarely, the logical count
``m``).  The driver below is ``Frugal.run`` (``_schemes.py:566-649``) restated
around the reference's own methods with that one loop in ascending edge id; the
unmodified ``Frugal.run`` is executed on the same seed next to it and its
``m`` / ``step_counts`` are stored too (``unmodified_*``), so the tests can see
how often the iteration order matters.
"""
from __future__ import annotations

import itertools
import math
import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.decoders import Snowflake  # noqa: E402
from localuf.decoders.snowflake.main import FloorContact  # noqa: E402

SF_CASES = [
    # name, spec, decoder kwargs, p, n, seed
    ('sf_rep3_phen', dict(code='Repetition', d=3, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.08, 40, 51),
    ('sf_rep5_phen', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.10, 40, 52),
    ('sf_rep5_phen_11', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.10, 40, 53),
    ('sf_rep5_phen_slow', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(merger='slow'), 0.10, 30, 54),
    ('sf_sf3_phen', dict(code='Surface', d=3, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.04, 40, 55),
    ('sf_sf5_phen', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.03, 12, 56),
    ('sf_sf5_phen_11', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.03, 10, 57),
    ('sf_sf3_cl', dict(code='Surface', d=3, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.02, 40, 58),
    ('sf_sf5_cl', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.010, 12, 59),
    ('sf_sf5_cl_slow11', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
     dict(merger='slow', schedule='1:1'), 0.010, 8, 60),
    ('sf_sf7_cl', dict(code='Surface', d=7, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.006, 4, 61),
    ('sf_sf5_phen_b2', dict(code='Surface', d=5, noise='phenomenological',
                            kwargs=dict(scheme='frugal', buffer_height=2)), {}, 0.02, 10, 62),
]


def _phases(scheme, d, p, n):
    h, c = scheme.WINDOW_HEIGHT, scheme._COMMIT_HEIGHT
    return itertools.chain(
        ((p, math.ceil(h / c), False, False),), itertools.repeat((p, d, True, False), n),
        ((p, d - 1, False, False),), ((p, 1, False, True),), ((0, 2 * h, False, False),))


def _drive(code, dec, p, n, seed, time_only, eidx):
    """Frugal.run with the floor error edges loaded in ascending edge id."""
    import gen_golden as G
    scheme = code.SCHEME
    E = len(code.EDGES)
    floor = [e for e, edge in dec.EDGES.items() if isinstance(edge.CONTACT, FloorContact)]
    random.seed(seed)
    scheme.reset(); dec.reset()
    rec = dict(fresh=[], excl=[], runtime=[], floor=[], m=[], timed=[])
    for p_inst, count, timed, excl in _phases(scheme, code.D, p, n):
        for _ in range(count):
            # Frugal._raise_window, canonical iteration order
            nxt = set()
            for e in sorted(scheme.error, key=eidx.__getitem__):
                if any(v[code.TIME_AXIS] == 0 for v in e):
                    scheme.pairs.load(e)
                else:
                    nxt.add(code.raise_edge(e, delta_t=-scheme._COMMIT_HEIGHT))
            scheme.error = nxt
            error = code.make_error(p_inst, exclude_future_boundary=excl)
            syndrome = scheme._load(error)
            rec['fresh'].append(G.pack((eidx[e] for e in error), E))
            rec['excl'].append(int(excl))
            rec['floor'].append(G.pack((eidx[e] for e in floor if dec.EDGES[e].correction), E))
            rec['runtime'].append(dec.decode(syndrome, time_only=time_only))
            rec['m'].append(scheme.get_logical_error())
            rec['timed'].append(int(timed))
    return rec


def sf_case(name, spec, dec_kw, p, n, seed):
    import gen_golden as G
    code = G.make_code(spec)
    assert code.SCHEME._COMMIT_HEIGHT == 1
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    dec = Snowflake(code, **dec_kw)
    a = _drive(code, dec, p, n, seed, 'all', eidx)
    u = _drive(code, dec, p, n, seed, 'unrooting', eidx)
    assert all(np.array_equal(x, y) for x, y in zip(a['fresh'], u['fresh']))
    # the unmodified scheme loop on the same seed
    random.seed(seed)
    m_unmod, slenderness = code.SCHEME.run(dec, p, n, time_only='merging')
    steps_unmod = list(code.SCHEME.step_counts)
    loops = 1 if dec_kw.get('schedule') == '1:1' else 2
    merging = np.array(a['runtime']) - 2 * loops
    timed = np.array(a['timed'], bool)
    d = code.D
    assert steps_unmod == [int(x) for x in merging[timed].reshape(n, d).sum(axis=1)]
    arrays = dict(fresh_bits=np.array(a['fresh']), excl=np.array(a['excl'], np.uint8),
                  runtime_all=np.array(a['runtime'], np.int32), runtime_unrooting=np.array(u['runtime'], np.int32),
                  floor_bits=np.array(a['floor']), m_cycle=np.array(a['m'], np.int32), timed=timed.astype(np.uint8),
                  unmodified_steps=np.array(steps_unmod, np.int32))
    meta = dict(kind='snowflake', spec=spec, decoder_kwargs=dec_kw, p=p, n=n, seed=seed, E=len(code.EDGES),
                cycles=len(a['runtime']), m=int(sum(a['m'])), unmodified_m=int(m_unmod), slenderness=slenderness,
                source='unmodified reference Snowflake driven by Frugal.run restated with the floor error '
                       'edges loaded in ascending edge id')
    return meta, arrays


def jobs():
    return {case[0]: (lambda c=case: sf_case(*c)) for case in SF_CASES}
