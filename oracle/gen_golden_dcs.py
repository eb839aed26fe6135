"""Golden vectors for the decoder confidence scores, exported from the UNMODIFIED reference.
TEST INFRASTRUCTURE ONLY; imported by oracle/gen_golden.py.

``dcs_*``: growth states of the reference ``UF`` (static: the growth does not depend on the merge order;
dynamic: the canonical-order subclass, for BURNT edges) with ``swim_distance``, ``unclustered_edge_fraction``
(``localuf/decoders/_base_uf.py:106-121,250-316``) at ``noise_level`` None and p, and
``unclustered_node_fraction`` (``uf.py:188-197``).

``dcsf_*``: one ``Frugal.run`` of the reference ``Snowflake`` (driver of oracle/gen_golden_sf.py) with, after
every decoding cycle, the scores ``Snowflake.decode(metrics=...)`` records (``snowflake/main.py:266-297``).
"""
from __future__ import annotations

import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.constants import Growth  # noqa: E402
from localuf.decoders import UF, Snowflake  # noqa: E402
from localuf.decoders.snowflake.main import FloorContact  # noqa: E402
from ref_canonical import CanonicalUF  # noqa: E402

DCS_CASES = [
    ('dcs_sf5_cc_p10', dict(code='Surface', d=5, noise='code capacity'), 0.10, 200, 71),
    ('dcs_sf9_cc_p08', dict(code='Surface', d=9, noise='code capacity'), 0.08, 100, 72),
    ('dcs_sf5_phen_p04', dict(code='Surface', d=5, noise='phenomenological'), 0.04, 100, 73),
    ('dcs_sf7_phen_p02', dict(code='Surface', d=7, noise='phenomenological'), 0.02, 40, 74),
    ('dcs_sf5_cl_p015', dict(code='Surface', d=5, noise='circuit-level'), 0.015, 60, 75),
    ('dcs_rep7_phen_p08', dict(code='Repetition', d=7, noise='phenomenological'), 0.08, 100, 76),
]


def _codes(dec, code):
    return np.array([3 if dec.growth[e] is Growth.BURNT else int(dec.growth[e]) for e in code.EDGES], np.uint8)


def dcs_case(name, spec, p, shots, seed):
    import gen_golden as G
    code = G.make_code(spec)
    ids, eidx, N, B, eu, ev = G.flatten(code)
    E, D = len(code.EDGES), N - B
    variants = {'static': UF(code), 'dynamic': CanonicalUF(code, dynamic=True)}
    weights = code.NOISE.get_edge_weights(p)
    arrays = dict(syn_bits=np.zeros((shots, G.words(D)), np.uint32),
                  weight_p=np.array([weights[e][1] for e in code.EDGES], np.float64))
    for k in variants:
        arrays[f'{k}_growth'] = np.zeros((shots, E), np.uint8)
        for f in ('swim_none', 'swim_p', 'uef_none', 'uef_p', 'unf'):
            arrays[f'{k}_{f}'] = np.zeros(shots, np.float64)
    random.seed(seed)
    for s in range(shots):
        error = code.make_error(p)
        syndrome = code.get_syndrome(error)
        arrays['syn_bits'][s] = G.pack((ids[v] - B for v in syndrome), D)
        for k, dec in variants.items():
            dec.reset()
            dec.decode(set(syndrome))
            arrays[f'{k}_growth'][s] = _codes(dec, code)
            arrays[f'{k}_swim_none'][s] = dec.swim_distance()
            arrays[f'{k}_swim_p'][s] = dec.swim_distance(noise_level=p)
            arrays[f'{k}_uef_none'][s] = dec.unclustered_edge_fraction()
            arrays[f'{k}_uef_p'][s] = dec.unclustered_edge_fraction(noise_level=p)
            arrays[f'{k}_unf'][s] = dec.unclustered_node_fraction()
    meta = dict(kind='dcs', spec=spec, p=p, shots=shots, seed=seed, N=N, B=B, E=E,
                source='unmodified reference UF (static) / oracle/ref_canonical.CanonicalUF (dynamic): '
                       'swim_distance, unclustered_edge_fraction, unclustered_node_fraction after decode')
    return meta, arrays


DCSF_CASES = [
    ('dcsf_rep5_phen', dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.10, 12, 81),
    ('dcsf_sf3_phen', dict(code='Surface', d=3, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.04, 16, 82),
    ('dcsf_sf5_phen', dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), {}, 0.03, 6, 83),
    ('dcsf_sf3_cl', dict(code='Surface', d=3, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.02, 16, 84),
    ('dcsf_sf5_cl', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), {}, 0.010, 6, 85),
    ('dcsf_sf5_cl_11', dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
     dict(schedule='1:1'), 0.010, 5, 86),
]

METRICS = ('throughput', 'swim_distance', 'min_defect_height', 'unclustered_edge_fraction', 'min_active_layer')


def dcsf_case(name, spec, dec_kw, p, n, seed):
    import gen_golden as G
    from gen_golden_sf import _phases
    code = G.make_code(spec)
    scheme = code.SCHEME
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    E = len(code.EDGES)
    dec = Snowflake(code, **dec_kw)
    rec = dict(fresh=[], runtime=[], timed=[], swim_none=[], swim_p=[], uef_none=[], uef_p=[], mdh=[], mal=[], growth=[])
    random.seed(seed)
    scheme.reset(); dec.reset()
    for p_inst, count, timed, excl in _phases(scheme, code.D, p, n):
        for _ in range(count):
            nxt = set()                                   # Frugal._raise_window, ascending edge id
            for e in sorted(scheme.error, key=eidx.__getitem__):
                if any(v[code.TIME_AXIS] == 0 for v in e):
                    scheme.pairs.load(e)
                else:
                    nxt.add(code.raise_edge(e, delta_t=-scheme._COMMIT_HEIGHT))
            scheme.error = nxt
            error = code.make_error(p_inst, exclude_future_boundary=excl)
            syndrome = scheme._load(error)
            rec['fresh'].append(G.pack((eidx[e] for e in error), E))
            before = {k: len(v) for k, v in dec.confidence_score_history.items()}
            mal_before = sum(dec.min_active_layers.values())
            rec['runtime'].append(dec.decode(syndrome, time_only='all', metrics=METRICS, noise_level_for_priors=p))
            scheme.get_logical_error()
            h = dec.confidence_score_history
            assert all(len(h[k]) == before.get(k, 0) + 1 for k in METRICS if k != 'min_active_layer')
            assert sum(dec.min_active_layers.values()) == mal_before + 1
            assert h['throughput'][-1] == 1 / rec['runtime'][-1]
            rec['swim_p'].append(h['swim_distance'][-1])
            rec['uef_p'].append(h['unclustered_edge_fraction'][-1])
            rec['mdh'].append(h['min_defect_height'][-1])
            rec['mal'].append(dec.min_active_layer())
            rec['swim_none'].append(dec.swim_distance())
            rec['uef_none'].append(dec.unclustered_edge_fraction())
            rec['timed'].append(int(timed))
            g = dec.growth
            rec['growth'].append(np.array([int(g[e]) if e in g else 0 for e in code.EDGES], np.uint8))
    # the unmodified Frugal.run on the same seed: its histories are the timed cycles of the record above
    random.seed(seed)
    code.SCHEME.run(dec, p, n, time_only='all', metrics=METRICS)
    timed = np.array(rec['timed'], bool)
    hist = dec.confidence_score_history
    unmod = dict(swim_p=np.array(hist['swim_distance']), uef_p=np.array(hist['unclustered_edge_fraction']),
                 mdh=np.array(hist['min_defect_height']), throughput=np.array(hist['throughput']))
    same = bool(np.array_equal(unmod['swim_p'], np.array(rec['swim_p'])[timed])
                and np.array_equal(unmod['mdh'], np.array(rec['mdh'])[timed]))
    weights = code.NOISE.get_edge_weights(p)
    arrays = dict(fresh_bits=np.array(rec['fresh']), runtime_all=np.array(rec['runtime'], np.int32),
                  timed=timed.astype(np.uint8), growth=np.array(rec['growth']),
                  swim_none=np.array(rec['swim_none'], np.float64), swim_p=np.array(rec['swim_p'], np.float64),
                  uef_none=np.array(rec['uef_none'], np.float64), uef_p=np.array(rec['uef_p'], np.float64),
                  min_defect_height=np.array(rec['mdh'], np.int32), min_active_layer=np.array(rec['mal'], np.int32),
                  weight_p=np.array([weights[e][1] for e in code.EDGES], np.float64),
                  in_growth=np.array([e in dec.growth for e in code.EDGES], np.uint8),
                  unmodified_min_active_layers=np.array(sorted(dec.min_active_layers.items()), np.int64).reshape(-1, 2))
    for k, v in unmod.items():
        arrays[f'unmodified_{k}'] = v
    meta = dict(kind='dcs snowflake', spec=spec, decoder_kwargs=dec_kw, p=p, n=n, seed=seed, E=E,
                cycles=len(rec['runtime']), unmodified_run_same_stream=same,
                source='unmodified reference Snowflake.decode(metrics=...) driven by Frugal.run restated with the '
                       'floor error edges loaded in ascending edge id; the unmodified Frugal.run(metrics=...) next to it')
    return meta, arrays


def jobs():
    out = {case[0]: (lambda c=case: dcs_case(*c)) for case in DCS_CASES}
    out.update({case[0]: (lambda c=case: dcsf_case(*c)) for case in DCSF_CASES})
    return out
