"""C oracle of Macar / Actis (oracle/luf_oracle_ca.c) vs. golden vectors
exported from the unmodified reference (tests/golden/ca_*.npz)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

DECODERS = {'macar': (1, 0), 'actis': (2, 0), 'actis_unopt': (2, 4)}


@pytest.mark.parametrize('name', golden_names('ca_'))
def test_ca_oracle_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    orc = binding.Oracle.from_code(code)
    for key in meta['decoders']:
        dec, flags = DECODERS[key]
        corr, growth, steps = orc.decode_batch(dec, flags, arr['syn_bits'])
        np.testing.assert_array_equal(steps, arr[f'{key}_steps'], err_msg=f'{name} {key} (tSV, tBP)')
        np.testing.assert_array_equal(growth, arr[f'{key}_growth'], err_msg=f'{name} {key} erasure after validate')
        np.testing.assert_array_equal(corr, arr[f'{key}_corr'], err_msg=f'{name} {key} correction')
        if key == 'macar' and str(code.SCHEME) == 'batch':
            np.testing.assert_array_equal(orc.logical_batch(arr['err_bits'], corr), arr['macar_logical'])


def test_ca_known_answers():
    """SURVEY.md appendix A3 probes / reference tests
    (tests/decoders/luf/main/LUF/test_LUF_integration.py): Surface d=5."""
    import localuf_b200 as L
    for noise, macar_empty, actis_empty in (('code capacity', (4, 2), (35, 22)),
                                            ('phenomenological', (4, 2), (47, 30)),
                                            ('circuit-level', (4, 2), (23, 14))):
        code = L.Surface(5, noise)
        orc = binding.Oracle.from_code(code)
        empty = np.zeros((1, orc.WD), np.uint32)
        assert tuple(orc.decode_batch(1, 0, empty)[2][0]) == macar_empty
        assert tuple(orc.decode_batch(2, 0, empty)[2][0]) == actis_empty
    code = L.Surface(5, 'code capacity')
    g = code.FLAT
    orc = binding.Oracle.from_code(code)
    corr, _, steps = orc.decode_batch(1, 0, g.syndrome_to_bits({(2, 1)})[None, :])
    assert tuple(steps[0]) == (30, 6)
    assert g.bits_to_edges(corr[0]) == {((2, -1), (2, 0)), ((2, 0), (2, 1))}
    _, _, steps = orc.decode_batch(1, 0, g.syndrome_to_bits({(2, 1), (2, 2)})[None, :])
    assert tuple(steps[0]) == (6, 3)
