"""Golden vectors for NodeUF (frontier-node growth, localuf/decoders/node_uf.py),
exported from the reference.  TEST INFRASTRUCTURE ONLY; imported by
oracle/gen_golden.py.  Same layout as the ``uf_*`` cases; growth of the
canonical execution is asserted equal to the unmodified ``NodeUF`` on every shot."""
from __future__ import annotations

import random

import numpy as np

import ref_shim

localuf = ref_shim.load()
from localuf.constants import Growth  # noqa: E402
from localuf.decoders import NodeUF, NodeBUF, BUF  # noqa: E402
from ref_canonical import CanonicalNodeUF, CanonicalNodeBUF, CanonicalBUF  # noqa: E402

NUF_CASES = [
    ('nuf_sf7_cc_p12', dict(code='Surface', d=7, noise='code capacity'), 0.12, 300, 31),
    ('nuf_sf9_cc_p08', dict(code='Surface', d=9, noise='code capacity'), 0.08, 300, 32),
    ('nuf_sf5_phen_p05', dict(code='Surface', d=5, noise='phenomenological'), 0.05, 200, 33),
    ('nuf_sf11_phen_p015', dict(code='Surface', d=11, noise='phenomenological'), 0.015, 50, 34),
    ('nuf_sf5_cl_p02', dict(code='Surface', d=5, noise='circuit-level'), 0.02, 150, 35),
    ('nuf_rep7_phen_p06', dict(code='Repetition', d=7, noise='phenomenological'), 0.06, 200, 36),
    ('nuf_sf4_cl_fwd_p03', dict(code='Surface', d=4, noise='circuit-level', kwargs=dict(scheme='forward')), 0.03, 150, 37),
]


def nuf_case(name, spec, p, shots, seed, canonical=CanonicalNodeUF, unmodified=NodeUF):
    import gen_golden as G
    code = G.make_code(spec)
    ids, eidx, N, B, eu, ev = G.flatten(code)
    E, D = len(code.EDGES), N - B
    variants = {'default': canonical(code), 'west': canonical(code, inclination='west'),
                'dynamic': canonical(code, dynamic=True)}
    plain = unmodified(code)
    is_batch = str(code.SCHEME) == 'batch'
    err_bits = np.zeros((shots, G.words(E)), np.uint32)
    syn_bits = np.zeros((shots, G.words(D)), np.uint32)
    out = {k: dict(growth=np.zeros((shots, E), np.uint8), corr=np.zeros((shots, G.words(E)), np.uint32),
                   logical=np.zeros(shots, np.uint8), rounds=np.zeros(shots, np.int32)) for k in variants}
    raised = np.zeros(shots, np.uint8)
    random.seed(seed)
    for s in range(shots):
        error = code.make_error(p) if is_batch else {e for e in code.EDGES if random.random() < p}
        syndrome = code.get_syndrome(error)
        err_bits[s] = G.pack((eidx[e] for e in error), E)
        syn_bits[s] = G.pack((ids[v] - B for v in syndrome), D)
        for k, dec in variants.items():
            dec.reset()
            dec.decode(set(syndrome))
            o = out[k]
            o['growth'][s] = np.array([3 if dec.growth[e] is Growth.BURNT else int(dec.growth[e]) for e in code.EDGES], np.uint8)
            o['corr'][s] = G.pack((eidx[e] for e in dec.correction), E)
            o['rounds'][s] = dec.round_count
            assert code.get_syndrome(dec.correction) == syndrome
            if is_batch:
                o['logical'][s] = code.get_logical_error(error ^ dec.correction)
        # the unmodified decoder; BUF raises ValueError (Growth(3), constants.py:17-31) when an edge that one
        # cluster grew to HALF is later grown by both its clusters in one round -- those shots are
        # flagged, the canonical execution saturates at FULL there
        try:
            plain.reset(); plain.decode(set(syndrome))
        except ValueError:
            raised[s] = 1
        else:
            assert all(plain.growth[e] == variants['default'].growth[e] for e in code.EDGES)
            assert code.get_syndrome(plain.correction) == syndrome
    arrays = dict(err_bits=err_bits, syn_bits=syn_bits, unmodified_raised=raised)
    for k, o in out.items():
        for f, a in o.items():
            arrays[f'{k}_{f}'] = a
    meta = dict(kind=unmodified.__name__, spec=spec, p=p, shots=shots, seed=seed, N=N, B=B, E=E,
                source=f'reference {unmodified.__name__} via oracle/ref_canonical.{canonical.__name__}; '
                       f'growth asserted equal to the unmodified decoder')
    return meta, arrays


# NodeBUF keys its buckets by cluster size up to (d+1)*d (node_buf.py:26-28) and raises KeyError
# beyond (already on Surface d=3 with two measurement rounds): code capacity only.
NBUF_CASES = [
    ('nbuf_sf7_cc_p12', dict(code='Surface', d=7, noise='code capacity'), 0.12, 300, 41),
    ('nbuf_sf9_cc_p08', dict(code='Surface', d=9, noise='code capacity'), 0.08, 300, 42),
    ('nbuf_sf9_cc_p15', dict(code='Surface', d=9, noise='code capacity'), 0.15, 200, 43),
    ('nbuf_sf5_cc_p20', dict(code='Surface', d=5, noise='code capacity'), 0.20, 300, 44),
    ('nbuf_sf3_cc_p15', dict(code='Surface', d=3, noise='code capacity'), 0.15, 300, 45),
]


BUF_CASES = [
    ('buf_sf7_cc_p12', dict(code='Surface', d=7, noise='code capacity'), 0.12, 300, 51),
    ('buf_sf9_cc_p08', dict(code='Surface', d=9, noise='code capacity'), 0.08, 300, 52),
    ('buf_sf5_phen_p05', dict(code='Surface', d=5, noise='phenomenological'), 0.05, 200, 53),
    ('buf_sf11_phen_p015', dict(code='Surface', d=11, noise='phenomenological'), 0.015, 50, 54),
    ('buf_sf5_cl_p02', dict(code='Surface', d=5, noise='circuit-level'), 0.02, 150, 55),
    ('buf_rep7_phen_p06', dict(code='Repetition', d=7, noise='phenomenological'), 0.06, 200, 56),
]


def jobs():
    out = {case[0]: (lambda c=case: nuf_case(*c)) for case in NUF_CASES}
    out.update({case[0]: (lambda c=case: nuf_case(*c, canonical=CanonicalNodeBUF, unmodified=NodeBUF)) for case in NBUF_CASES})
    out.update({case[0]: (lambda c=case: nuf_case(*c, canonical=CanonicalBUF, unmodified=BUF)) for case in BUF_CASES})
    return out
