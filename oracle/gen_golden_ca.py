"""Golden vectors for the cellular-automaton decoders (Macar, Actis), exported
from the UNMODIFIED reference.  TEST INFRASTRUCTURE ONLY; imported by
oracle/gen_golden.py.  These decoders are fully deterministic in the reference
(finding 1), so no canonical-order shim is involved.
"""
from __future__ import annotations

import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.constants import Growth  # noqa: E402
from localuf.decoders import Macar, Actis  # noqa: E402

CA_CASES = [
    # name, spec, p, shots, seed, with_actis
    ('ca_sf3_cc_p10', dict(code='Surface', d=3, noise='code capacity'), 0.10, 120, 31, True),
    ('ca_sf5_cc_p08', dict(code='Surface', d=5, noise='code capacity'), 0.08, 120, 32, True),
    ('ca_sf7_cc_p10', dict(code='Surface', d=7, noise='code capacity'), 0.10, 80, 33, True),
    ('ca_sf9_cc_p05', dict(code='Surface', d=9, noise='code capacity'), 0.05, 60, 34, False),
    ('ca_sf3_phen_p05', dict(code='Surface', d=3, noise='phenomenological'), 0.05, 100, 35, True),
    ('ca_sf5_phen_p03', dict(code='Surface', d=5, noise='phenomenological'), 0.03, 60, 36, True),
    ('ca_sf7_phen_p02', dict(code='Surface', d=7, noise='phenomenological'), 0.02, 24, 37, False),
    ('ca_sf3_cl_p02', dict(code='Surface', d=3, noise='circuit-level'), 0.02, 80, 38, True),
    ('ca_sf5_cl_p01', dict(code='Surface', d=5, noise='circuit-level'), 0.01, 30, 39, True),
    ('ca_sf4_phen_fwd_p03', dict(code='Surface', d=4, noise='phenomenological',
                                 kwargs=dict(scheme='forward')), 0.03, 40, 40, False),
    ('ca_sf11_phen_p01', dict(code='Surface', d=11, noise='phenomenological'), 0.01, 6, 41, False),
]


def _growth_codes(dec, code):
    return np.array([3 if dec.growth[e] is Growth.BURNT else int(dec.growth[e]) for e in code.EDGES], np.uint8)


def ca_case(name, spec, p, shots, seed, with_actis):
    import gen_golden as G
    code = G.make_code(spec)
    ids, eidx, N, B, eu, ev = G.flatten(code)
    E, D = len(code.EDGES), N - B
    is_batch = str(code.SCHEME) == 'batch'
    decs = {'macar': Macar(code)}
    if with_actis:
        decs['actis'] = Actis(code)
        decs['actis_unopt'] = Actis(code, _optimal=False)
    err_bits = np.zeros((shots, G.words(E)), np.uint32)
    syn_bits = np.zeros((shots, G.words(D)), np.uint32)
    out = {k: dict(growth=np.zeros((shots, E), np.uint8), corr=np.zeros((shots, G.words(E)), np.uint32),
                   steps=np.zeros((shots, 2), np.int32), logical=np.zeros(shots, np.uint8)) for k in decs}
    random.seed(seed)
    for s in range(shots):
        if s == 0:
            error = set()                      # the empty syndrome is a known answer of the reference tests
        elif is_batch:
            error = code.make_error(p)
        else:
            error = {e for e in code.EDGES if random.random() < p}
        syndrome = code.get_syndrome(error)
        err_bits[s] = G.pack((eidx[e] for e in error), E)
        syn_bits[s] = G.pack((ids[v] - B for v in syndrome), D)
        for k, dec in decs.items():
            dec.reset()
            tSV = dec.validate(set(syndrome))
            out[k]['growth'][s] = _growth_codes(dec, code)      # erasure after syndrome validation
            tBP = dec.peel(False)
            out[k]['steps'][s] = (tSV, tBP)
            out[k]['corr'][s] = G.pack((eidx[e] for e in dec.correction), E)
            if k == 'macar':
                assert code.get_syndrome(dec.correction) == syndrome
                if is_batch:
                    out[k]['logical'][s] = code.get_logical_error(error ^ dec.correction)
            else:
                assert dec.correction == set()
    arrays = dict(err_bits=err_bits, syn_bits=syn_bits, edge_u=eu, edge_v=ev)
    for k, o in out.items():
        for f, a in o.items():
            arrays[f'{k}_{f}'] = a
    meta = dict(kind='ca', spec=spec, p=p, shots=shots, seed=seed, N=N, B=B, E=E, decoders=list(decs),
                source='unmodified reference Macar / Actis (validate, then peel)')
    return meta, arrays


def jobs():
    return {case[0]: (lambda c=case: ca_case(*c)) for case in CA_CASES}
