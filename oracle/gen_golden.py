"""Export golden vectors from the real reference into ``tests/golden/``.

TEST INFRASTRUCTURE ONLY.  Run in the authoring container, where the upstream
reference is mounted read-only at ``/root/reference``:

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden.py [--only NAME ...]

Everything written here comes from *reference* objects (graphs from the
reference ``Surface``/``Repetition``; errors from the reference samplers with
``random.seed``; decoder outputs from the reference decoders, UF through the
canonical-order subclass of ``oracle/ref_canonical.py``).  None of the product
package is imported, so the fixtures are an independent witness for the C
oracle and for the CUDA kernels.

Flat conventions (identical to ``localuf_b200/_graph.py``): node index = rank
under (boundary first, reference ``index_to_id``) -- equal to ``index_to_id``
itself for every graph except the 1-D repetition code; edge index = position in ``code.EDGES``; bit ``k`` of
little-endian 32-bit word ``w`` is element ``32 w + k``; detector bit ``k`` is
node ``B + k`` with ``B`` the number of boundary nodes.
"""
from __future__ import annotations

import argparse
import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

localuf = ref_shim.load()
from localuf import Repetition, Surface  # noqa: E402
from localuf.constants import Growth  # noqa: E402
from localuf.decoders import UF, Macar, Actis  # noqa: E402
from ref_canonical import CanonicalUF  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')

CODES = {'Surface': Surface, 'Repetition': Repetition}


def words(n):
    return (n + 31) // 32


def pack(indices, nbits):
    out = np.zeros(words(nbits), dtype=np.uint32)
    for k in indices:
        out[k >> 5] |= np.uint32(1 << (k & 31))
    return out


def flatten(code):
    """Reference graph -> flat arrays keyed by reference ids."""
    order = sorted(code.NODES, key=lambda v: (not code.is_boundary(v), code.index_to_id(v), v))
    ids = {v: k for k, v in enumerate(order)}     # boundary nodes first, then by reference id
    N = len(ids)
    B = len(code.BOUNDARY_NODES)
    if len(code.NODES[0]) > 1 and not code.MERGED_EQUIVALENT_BOUNDARY_NODES:
        # everywhere except the 1-D repetition graph this IS the reference id
        assert all(ids[v] == code.index_to_id(v) for v in code.NODES)
    eu = np.array([ids[e[0]] for e in code.EDGES], dtype=np.int32)
    ev = np.array([ids[e[1]] for e in code.EDGES], dtype=np.int32)
    eidx = {e: k for k, e in enumerate(code.EDGES)}
    return ids, eidx, N, B, eu, ev


def make_code(spec):
    kw = dict(spec.get('kwargs', {}))
    return CODES[spec['code']](spec['d'], spec['noise'], **kw)


# --------------------------------------------------------------------------- UF
UF_CASES = [
    # name, code spec, p, shots, seed
    ('uf_rep5_cc_p05', dict(code='Repetition', d=5, noise='code capacity'), 0.05, 400, 11),
    ('uf_rep5_cc_p30', dict(code='Repetition', d=5, noise='code capacity'), 0.30, 300, 12),
    ('uf_rep7_phen_p05', dict(code='Repetition', d=7, noise='phenomenological'), 0.05, 300, 13),
    ('uf_sf3_cc_p10', dict(code='Surface', d=3, noise='code capacity'), 0.10, 300, 14),
    ('uf_sf7_cc_p25', dict(code='Surface', d=7, noise='code capacity'), 0.25, 300, 15),
    ('uf_sf9_cc_p05', dict(code='Surface', d=9, noise='code capacity'), 0.05, 400, 16),
    ('uf_sf9_cc_p10', dict(code='Surface', d=9, noise='code capacity'), 0.10, 400, 17),
    ('uf_sf5_phen_p03', dict(code='Surface', d=5, noise='phenomenological'), 0.03, 300, 18),
    ('uf_sf5_phen_p08', dict(code='Surface', d=5, noise='phenomenological'), 0.08, 200, 19),
    ('uf_sf11_phen_p01', dict(code='Surface', d=11, noise='phenomenological'), 0.01, 60, 20),
    ('uf_sf11_phen_p026', dict(code='Surface', d=11, noise='phenomenological'), 0.026, 60, 21),
    ('uf_sf5_cl_p01', dict(code='Surface', d=5, noise='circuit-level'), 0.01, 200, 22),
    ('uf_sf5_cl_p03', dict(code='Surface', d=5, noise='circuit-level'), 0.03, 100, 23),
    ('uf_sf4_phen_fwd_p04', dict(code='Surface', d=4, noise='phenomenological',
                                 kwargs=dict(scheme='forward')), 0.04, 150, 24),
    # circuit-level forward window: future-boundary nodes have degree > 1 (interior nodes of the peeling tree)
    ('uf_sf4_cl_fwd_p03', dict(code='Surface', d=4, noise='circuit-level',
                               kwargs=dict(scheme='forward')), 0.03, 200, 25),
    ('uf_sf5_cl_fwd_p015', dict(code='Surface', d=5, noise='circuit-level',
                                kwargs=dict(scheme='forward', commit_height=2, buffer_height=3)), 0.015, 150, 26),
]


def uf_case(name, spec, p, shots, seed):
    code = make_code(spec)
    ids, eidx, N, B, eu, ev = flatten(code)
    E, D = len(code.EDGES), N - B
    variants = {
        'default': CanonicalUF(code),
        'west': CanonicalUF(code, inclination='west'),
        'dynamic': CanonicalUF(code, dynamic=True),
    }
    plain = UF(code)                         # unmodified reference
    plain_west = UF(code, inclination='west')
    is_batch = str(code.SCHEME) == 'batch'

    err_bits = np.zeros((shots, words(E)), np.uint32)
    syn_bits = np.zeros((shots, words(D)), np.uint32)
    out = {k: dict(growth=np.zeros((shots, E), np.uint8),
                   corr=np.zeros((shots, words(E)), np.uint32),
                   logical=np.zeros(shots, np.uint8),
                   rounds=np.zeros(shots, np.int32)) for k in variants}
    west_logical_unmodified = np.zeros(shots, np.uint8)

    random.seed(seed)
    for s in range(shots):
        # streaming windows: sample every edge of the window directly (their own
        # samplers only draw the fresh layers)
        if is_batch:     # the reference's own sampler (also pins host make_error under random.seed)
            error = code.make_error(p)
        else:
            error = {e for e in code.EDGES if random.random() < p}
        syndrome = code.get_syndrome(error)
        err_bits[s] = pack((eidx[e] for e in error), E)
        syn_bits[s] = pack((ids[v] - B for v in syndrome), D)
        for k, dec in variants.items():
            dec.reset()
            dec.decode(set(syndrome))
            g = np.array([3 if dec.growth[e] is Growth.BURNT else int(dec.growth[e])
                          for e in code.EDGES], np.uint8)
            o = out[k]
            o['growth'][s] = g
            o['corr'][s] = pack((eidx[e] for e in dec.correction), E)
            o['rounds'][s] = dec.round_count
            assert code.get_syndrome(dec.correction) == syndrome
            if is_batch:
                o['logical'][s] = code.get_logical_error(error ^ dec.correction)
        # cross-checks against the UNMODIFIED reference
        plain.reset(); plain.decode(set(syndrome))
        assert all(plain.growth[e] == variants['default'].growth[e] for e in code.EDGES), \
            "growth of the canonical execution must equal the unmodified reference"
        assert code.get_syndrome(plain.correction) == syndrome
        if is_batch:
            plain_west.reset(); plain_west.decode(set(syndrome))
            west_logical_unmodified[s] = code.get_logical_error(error ^ plain_west.correction)
            assert west_logical_unmodified[s] == out['west']['logical'][s], \
                "west inclination: logical outcome is order-independent"
    arrays = dict(err_bits=err_bits, syn_bits=syn_bits, edge_u=eu, edge_v=ev,
                  west_logical_unmodified=west_logical_unmodified)
    for k, o in out.items():
        for f, a in o.items():
            arrays[f'{k}_{f}'] = a
    meta = dict(kind='uf', spec=spec, p=p, shots=shots, seed=seed, N=N, B=B, E=E,
                source='reference UF via oracle/ref_canonical.CanonicalUF; '
                       'growth asserted equal to unmodified reference UF')
    return meta, arrays


# ------------------------------------------------------------------------ graphs
GRAPH_SPECS = [
    dict(code='Repetition', d=5, noise='code capacity'),
    dict(code='Repetition', d=4, noise='phenomenological'),
    dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='forward')),
    dict(code='Repetition', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
    dict(code='Surface', d=3, noise='code capacity'),
    dict(code='Surface', d=9, noise='code capacity'),
    dict(code='Surface', d=3, noise='phenomenological'),
    dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(window_height=3)),
    dict(code='Surface', d=11, noise='phenomenological'),
    dict(code='Surface', d=4, noise='phenomenological', kwargs=dict(scheme='forward')),
    dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')),
    dict(code='Surface', d=3, noise='circuit-level'),
    dict(code='Surface', d=5, noise='circuit-level'),
    dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='forward')),
    dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')),
    dict(code='Surface', d=13, noise='circuit-level', kwargs=dict(scheme='frugal')),
    dict(code='Surface', d=5, noise='circuit-level',
         kwargs=dict(parametrization='standard', demolition=True, monolingual=True)),
    dict(code='Surface', d=4, noise='circuit-level', kwargs=dict(scheme='global batch', window_height=8)),
]


def graph_cases():
    """Edge lists, ids, probability classes and class probabilities of the
    reference graphs (pins the host graph layer)."""
    arrays, metas = {}, []
    for k, spec in enumerate(GRAPH_SPECS):
        code = make_code(spec)
        ids, eidx, N, B, eu, ev = flatten(code)
        noise = code.NOISE
        fresh = noise._FRESH_EDGES
        if isinstance(fresh, dict):
            groups = list(fresh.values())
            probs = {p: list(noise._get_flip_probabilities(p).values()) for p in (0.001, 0.01)}
        else:
            groups = [fresh]
            probs = {p: [p] for p in (0.001, 0.01)}
        cls = np.full(len(code.EDGES), 255, np.uint8)
        for c, g in enumerate(groups):
            for e in g:
                cls[eidx[e]] = c
        coords = np.array([v for v, _ in sorted(ids.items(), key=lambda kv: kv[1])], np.int32)
        arrays[f'g{k}_edge_u'] = eu
        arrays[f'g{k}_edge_v'] = ev
        arrays[f'g{k}_edge_class'] = cls
        arrays[f'g{k}_coords'] = coords
        commit = getattr(code.SCHEME, '_COMMIT_EDGES', None)
        if commit is not None:
            arrays[f'g{k}_commit_edges'] = np.array(sorted(eidx[e] for e in commit), np.int32)
        metas.append(dict(spec=spec, N=N, B=B, E=len(code.EDGES), n_edges_attr=code.N_EDGES,
                          probs={str(p): v for p, v in probs.items()},
                          dimension=code.DIMENSION, height=code.SCHEME.WINDOW_HEIGHT))
    return dict(kind='graphs', cases=metas), arrays


def save(name, meta, arrays):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name + '.npz')
    np.savez_compressed(path, meta=np.array(json.dumps(meta)), **arrays)
    print(f'{name}: {os.path.getsize(path) / 1024:.1f} KiB')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', nargs='*')
    args = ap.parse_args()
    jobs = {'graphs': lambda: graph_cases()}
    for case in UF_CASES:
        jobs[case[0]] = (lambda c=case: uf_case(*c))
    for extra in ('gen_golden_ca', 'gen_golden_sf', 'gen_golden_fw', 'gen_golden_gl', 'gen_golden_subset', 'gen_golden_nuf', 'gen_golden_io', 'gen_golden_lat', 'gen_golden_dcs'):
        try:
            jobs.update(__import__(extra).jobs())
        except ImportError:
            pass
    for name, job in jobs.items():
        if args.only and name not in args.only:
            continue
        meta, arrays = job()
        save(name, meta, arrays)


if __name__ == '__main__':
    main()
