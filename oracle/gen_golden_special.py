"""Golden vectors for ``sim.accuracy.monte_carlo_special`` (localuf/sim/accuracy/monte_carlo.py:162-201):
the code's noise model is circuit-level, the decoder's graph phenomenological.

TEST INFRASTRUCTURE ONLY; run in the authoring container (reference at /root/reference):

    PYTHONDONTWRITEBYTECODE=1 python oracle/gen_golden_special.py

Per shot, with reference objects only: error = the circuit-level code's ``make_error(p)`` (``random.seed``),
then exactly what ``Batch._sim_cycle_given_error`` does (``_schemes.py:143-155``) with a decoder built on the
phenomenological code: ``get_syndrome`` on the circuit-level code, ``reset`` / ``decode``, leftover =
error ^ correction, ``get_logical_error``.  Decoders: the canonical-order UF of ref_canonical.py (default and
west inclination; for west the unmodified reference's outcome is stored next to it and asserted equal) and
the unmodified Macar.  Stored: error bits over the circuit-level edges, correction bits over the
phenomenological edges, the logical bit.
"""
from __future__ import annotations

import json
import os
import random
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

localuf = ref_shim.load()
from localuf import Surface  # noqa: E402
from localuf.decoders import UF, Macar  # noqa: E402
from ref_canonical import CanonicalUF  # noqa: E402
from gen_golden import flatten, pack, words  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden')
CASES = [('sp_sf3_p02', 3, 0.02, 300, 5), ('sp_sf5_p01', 5, 0.01, 300, 6), ('sp_sf7_p006', 7, 0.006, 120, 7)]


def main():
    for name, d, p, shots, seed in CASES:
        code = Surface(d, 'circuit-level')
        phen = Surface(d, 'phenomenological')
        ids_c, eidx_c, N, B, _, _ = flatten(code)
        ids_p, eidx_p, Np, Bp, _, _ = flatten(phen)
        assert ids_c == ids_p and (N, B) == (Np, Bp)
        decs = {'uf': CanonicalUF(phen), 'west': CanonicalUF(phen, inclination='west'),
                'west_unmodified': UF(phen, inclination='west'), 'macar': Macar(phen)}
        random.seed(seed)
        errors = [code.make_error(p) for _ in range(shots)]
        out = {'err_bits': np.stack([pack([eidx_c[e] for e in err], code.N_EDGES) for err in errors])}
        for key, dec in decs.items():
            corr, logical = [], []
            for err in errors:
                syndrome = code.get_syndrome(err)
                dec.reset()
                dec.decode(syndrome)
                leftover = err ^ dec.correction
                logical.append(code.SCHEME.get_logical_error(leftover))
                corr.append(pack([eidx_p[e] for e in dec.correction], phen.N_EDGES))
            out[f'{key}_logical'] = np.array(logical, dtype=np.uint8)
            if key != 'west_unmodified':
                out[f'{key}_corr'] = np.stack(corr)
        assert np.array_equal(out['west_logical'], out['west_unmodified_logical'])
        meta = dict(d=d, p=p, shots=shots, seed=seed, E_code=code.N_EDGES, E_decoder=phen.N_EDGES,
                    fails={k: int(out[f'{k}_logical'].sum()) for k in decs})
        np.savez_compressed(os.path.join(OUT, name + '.npz'), meta=json.dumps(meta), **out)
        print(name, meta['fails'])


if __name__ == '__main__':
    main()
