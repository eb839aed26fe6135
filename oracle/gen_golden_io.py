"""Golden vectors for Snowflake's bit-string I/O (``Snowflake.generate_output``,
``snowflake/main.py:656-726``), exported from the UNMODIFIED reference.  TEST
INFRASTRUCTURE ONLY; imported by oracle/gen_golden.py.

Per case: random syndrome vectors (one per decoding cycle) in, the reference's
``floor_history`` out, and -- from a second pass with ``Snowflake.decode`` called
cycle by cycle -- the runtime of every cycle for ``time_only`` = 'all' /
'merging' / 'unrooting'.
"""
from __future__ import annotations

import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.decoders import Snowflake  # noqa: E402

IO_CASES = [
    # name, spec, decoder kwargs, defect density, cycles, seed
    ('io_rep5_phen', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.20, 60, 71),
    ('io_rep5_phen_11', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.20, 60, 72),
    ('io_rep7_phen_notime', dict(code='Repetition', d=7, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(_include_timelike_lowest_edges=False), 0.15, 40, 73),
    ('io_sf3_phen', dict(code='Surface', d=3, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.10, 50, 74),
    ('io_sf5_phen_slow', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(merger='slow'), 0.05, 24, 75),
    ('io_sf3_cl', dict(code='Surface', d=3, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.10, 50, 76),
    ('io_sf5_cl_11', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.04, 24, 77),
    ('io_rep5_phen_order', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1', _neighbor_order=('U', 'W', 'E', 'D')), 0.20, 60, 79),          # the order of the demo notebook, cell 2
    ('io_sf5_cl_order', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
     dict(_neighbor_order=('SEU', 'S', 'SD', 'EU', 'E', 'U', 'D', 'W', 'WD', 'NU', 'N', 'NWD')), 0.05, 24, 80),
    ('io_sf5_phen_order', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
     dict(_neighbor_order=('U', 'D', 'E', 'W', 'S', 'N', 'NWD', 'NU', 'WD', 'EU', 'SD', 'SEU')), 0.06, 24, 81),
    ('io_sf5_cl_notime', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
     dict(_include_timelike_lowest_edges=False), 0.04, 16, 78),
]


def io_case(name, spec, dec_kw, density, cycles, seed):
    import gen_golden as G
    code = G.make_code(spec)
    dec = Snowflake(code, **dec_kw)
    length = dec._BITSTRING_CONVERTER.LENGTH
    rng = random.Random(seed)
    vectors = [''.join('1' if rng.random() < density else '0' for _ in range(length)) for _ in range(cycles)]
    dec.reset()
    floors = list(dec.generate_output(list(vectors)))
    h = code.SCHEME.WINDOW_HEIGHT
    assert len(floors) == cycles + h
    # the same stream given as coordinate sets must give the same output
    dec.reset()
    sets = dec._BITSTRING_CONVERTER.syndrome_vectors_to_sets(vectors)
    assert list(dec.generate_output([set(s) for s in sets])) == floors
    runtimes = {}
    for mode in ('all', 'merging', 'unrooting'):
        dec.reset()
        runtimes[mode] = [dec.decode(set(s), time_only=mode) for s in sets]
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    arrays = dict(
        vectors=np.array([[int(c) for c in v] for v in vectors], np.uint8),
        floors=np.array([[int(c) for c in f] for f in floors], np.uint8),
        lowest_edges=np.array([eidx[e.INDEX] for e in dec._LOWEST_EDGES], np.int32),
        runtime_all=np.array(runtimes['all'], np.int32), runtime_merging=np.array(runtimes['merging'], np.int32),
        runtime_unrooting=np.array(runtimes['unrooting'], np.int32))
    meta = dict(kind='snowflake bit-string io', spec=spec, decoder_kwargs=dec_kw, density=density, cycles=cycles,
                seed=seed, h=h, source='unmodified reference Snowflake.generate_output / Snowflake.decode')
    return meta, arrays


def jobs():
    out = {case[0]: (lambda c=case: io_case(*c)) for case in IO_CASES}
    out.update({case[0]: (lambda c=case: iod_case(*c)) for case in IOD_CASES})
    return out


# ---- decode(..., defects_possible=False): cycles on which the decoder only drops (snowflake/main.py:244-264)
IOD_CASES = [
    ('iod_rep5_phen', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.20, 60, 91),
    ('iod_sf3_cl', dict(code='Surface', d=3, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.10, 50, 92),
    ('iod_sf5_phen', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.05, 30, 93),
    ('iod_sf5_cl_11', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.04, 24, 94),
]


def iod_case(name, spec, dec_kw, density, cycles, seed):
    """Random syndrome vectors; on a random third of the cycles the decoder is told that no defect is possible
    (it then only drops, whatever the window holds).  Runtimes for the three time_only modes and the floor
    vectors logged by decode(log_floor_history=True)."""
    import gen_golden as G
    code = G.make_code(spec)
    dec = Snowflake(code, **dec_kw)
    conv = dec._BITSTRING_CONVERTER
    rng = random.Random(seed)
    vectors = [''.join('1' if rng.random() < density else '0' for _ in range(conv.LENGTH)) for _ in range(cycles)]
    drop_only = [rng.random() < 1 / 3 for _ in range(cycles)]
    sets = conv.syndrome_vectors_to_sets(vectors)
    runtimes, floors = {}, None
    for mode in ('all', 'merging', 'unrooting'):
        dec.reset()
        dec.init_history()
        runtimes[mode] = [dec.decode(set(s), time_only=mode, defects_possible=not d, log_floor_history=True)
                          for s, d in zip(sets, drop_only)]
        if floors is None:
            floors = list(dec.floor_history)
        assert floors == list(dec.floor_history)
    arrays = dict(
        vectors=np.array([[int(c) for c in v] for v in vectors], np.uint8),
        drop_only=np.array(drop_only, np.uint8),
        floors=np.array([[int(c) for c in f] for f in floors], np.uint8),
        runtime_all=np.array(runtimes['all'], np.int32), runtime_merging=np.array(runtimes['merging'], np.int32),
        runtime_unrooting=np.array(runtimes['unrooting'], np.int32))
    meta = dict(kind='snowflake decode with defects_possible', spec=spec, decoder_kwargs=dec_kw, density=density,
                cycles=cycles, seed=seed, source='unmodified reference Snowflake.decode(defects_possible=...)')
    return meta, arrays
