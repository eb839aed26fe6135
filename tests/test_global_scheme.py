"""Global batch scheme: one window of d*n layers,
logical error strings *counted* (localuf/_schemes.py:166-203).  Goldens:
oracle/gen_golden_gl.py (reference errors, canonical-UF / unmodified-Macar
corrections, the count of the leftover in ascending edge id)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))


def edges_of(code, bits):
    b = np.unpackbits(np.ascontiguousarray(bits, dtype=np.uint32).view(np.uint8), bitorder='little')
    return {code.EDGES[k] for k in np.flatnonzero(b[:code.N_EDGES]).tolist()}


@pytest.mark.parametrize('name', golden_names('gl_'))
def test_host_counting_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    scheme = code.SCHEME
    assert str(scheme) == 'global batch' and scheme.WINDOW_HEIGHT == code.D * meta['n']
    for key in ('uf', 'macar'):
        for err, corr, want in zip(arr['err_bits'], arr[f'{key}_corr'], arr[f'{key}_count']):
            scheme.reset()
            assert scheme.get_logical_error(edges_of(code, err ^ corr)) == want
            assert scheme.count_leftover_bits(err ^ corr) == want
    # the unmodified reference's own set order agrees on all but a handful of branching leftovers
    assert (arr['uf_count'] != arr['uf_count_unmodified']).mean() < 0.1


def test_run_requires_matching_height():
    import localuf_b200 as L
    code = L.Surface(3, 'phenomenological', scheme='global batch', window_height=6)
    with pytest.raises(AssertionError):
        code.SCHEME.run(L.decoders.UF(code), 0.01, 3)


@pytest.mark.gpu
@pytest.mark.parametrize('name', golden_names('gl_'))
def test_device_decode_and_count(name):
    import localuf_b200 as L
    from localuf_b200 import _engine
    import binding
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    syn = orc.syndrome(arr['err_bits'])
    for key, dec_id in (('uf', _engine.DEC_UF), ('macar', _engine.DEC_MACAR)):
        corr, _, steps = eng.decode_batch_host(dec_id, 0, syn)
        np.testing.assert_array_equal(corr, arr[f'{key}_corr'])
        if key == 'macar':
            np.testing.assert_array_equal(steps, arr['macar_steps'])
        counts = [code.SCHEME.count_leftover_bits(l) for l in arr['err_bits'] ^ corr]
        np.testing.assert_array_equal(counts, arr[f'{key}_count'])
        # the device counts the strings itself (luf_count_strings): per window, equal to the reference's count
        np.testing.assert_array_equal(eng.count_strings_host(arr['err_bits'], corr), arr[f'{key}_count'])
    # Global.run end to end: device-sampled windows, rate near the goldens' (loose 6 sigma bound)
    uf = L.decoders.UF(code)
    shots = 400
    m, slender = code.SCHEME.run(uf, meta['p'], meta['n'], shots=shots, seed=5)
    assert slender == meta['n'] * shots
    mean = arr['uf_count'].mean()
    assert abs(m / shots - mean) < 6 * (max(arr['uf_count'].var(), 0.05) * (1 / shots + 1 / len(arr['uf_count']))) ** 0.5 + 0.05
    # ... and the per-window counts of a device-sampled run equal the host mirror's on the same windows
    total, per = eng.mc_global(_engine.DEC_UF, 0, meta['p'], 5, 0, 300, chunk=128, want_counts=True)
    err, syn2 = eng.sample_host(meta['p'], 5, 0, 300)
    corr2, _, _ = eng.decode_batch_host(_engine.DEC_UF, 0, syn2, want_steps=False)
    np.testing.assert_array_equal(per, [code.SCHEME.count_leftover_bits(l) for l in err ^ corr2])
    assert total == int(per.sum())
    # sim.accuracy.monte_carlo takes the scheme by name, like the reference (window height = d * n)
    df = L.sim.accuracy.monte_carlo({code.D: [(meta['p'], meta['n'])]}, type(code), L.decoders.UF,
                                    meta['spec']['noise'], scheme='global batch')
    assert df.loc['n', (code.D, meta['p'])] == meta['n']
