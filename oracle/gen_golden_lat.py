"""Golden vectors for ``Frugal.sample_latency`` (``_schemes.py:651-698``), exported from the
UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY; imported by oracle/gen_golden.py.

Per case and seed: the errors the reference drew in each of the 2h + 1 decoding cycles (recorded by
wrapping ``code.make_error``, so the exclusion of future-boundary faults in the last noisy cycle is
already applied) and the latency it returned for ``time_only`` = 'all' / 'merging' / 'unrooting'
(three runs on the same seed = the same errors).
"""
from __future__ import annotations

import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.decoders import Snowflake  # noqa: E402

LAT_CASES = [
    # name, spec, decoder kwargs, p, samples
    ('lat_rep5_phen', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.10, 60),
    ('lat_rep7_phen_11', dict(code='Repetition', d=7, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.08, 40),
    ('lat_sf3_phen', dict(code='Surface', d=3, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.05, 60),
    ('lat_sf5_phen_slow', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(merger='slow'), 0.03, 30),
    ('lat_sf3_cl', dict(code='Surface', d=3, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.03, 60),
    ('lat_sf5_cl', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.012, 30),
    ('lat_sf5_cl_11', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.012, 30),
    ('lat_sf7_cl', dict(code='Surface', d=7, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.006, 10),
]


def lat_case(name, spec, dec_kw, p, samples):
    import gen_golden as G
    code = G.make_code(spec)
    dec = Snowflake(code, **dec_kw)
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    E = len(code.EDGES)
    h = code.SCHEME.WINDOW_HEIGHT
    drawn = []
    make_error = code.make_error

    def recording(*a, **k):
        error = make_error(*a, **k)
        drawn.append(G.pack((eidx[e] for e in error), E))
        return error
    code.make_error = recording
    fresh, lat = [], {m: [] for m in ('all', 'merging', 'unrooting')}
    for k in range(samples):
        rows = None
        for mode in lat:
            drawn.clear()
            random.seed(1000 + k)
            lat[mode].append(code.SCHEME.sample_latency(dec, p, time_only=mode))
            assert len(drawn) == 2 * h + 1
            if rows is None:
                rows = np.array(drawn)
            else:
                assert np.array_equal(rows, np.array(drawn))
        fresh.append(rows)
    code.make_error = make_error
    arrays = dict(fresh_bits=np.array(fresh), latency_all=np.array(lat['all'], np.int32),
                  latency_merging=np.array(lat['merging'], np.int32),
                  latency_unrooting=np.array(lat['unrooting'], np.int32))
    meta = dict(kind='snowflake latency', spec=spec, decoder_kwargs=dec_kw, p=p, samples=samples, E=E, h=h,
                source='unmodified reference Frugal.sample_latency, errors recorded from code.make_error')
    return meta, arrays


def jobs():
    return {case[0]: (lambda c=case: lat_case(*c)) for case in LAT_CASES}
