"""C oracle (oracle/luf_oracle.c) vs. golden vectors exported from the real
reference: this is what pins the oracle (task brief, section 3)."""
import os
import sys
from math import comb

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

FLAGS = {'default': 0, 'dynamic': 1, 'west': 2}


@pytest.mark.parametrize('name', golden_names('uf_'))
def test_uf_oracle_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    orc = binding.Oracle.from_code(code)
    np.testing.assert_array_equal(orc.syndrome(arr['err_bits']), arr['syn_bits'])
    for variant, flags in FLAGS.items():
        corr, growth, steps = orc.decode_batch(0, flags, arr['syn_bits'])
        np.testing.assert_array_equal(growth, arr[f'{variant}_growth'], err_msg=f'{name} {variant} growth')
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'], err_msg=f'{name} {variant} rounds')
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'], err_msg=f'{name} {variant} correction')
        if str(code.SCHEME) == 'batch':
            np.testing.assert_array_equal(orc.logical_batch(arr['err_bits'], corr), arr[f'{variant}_logical'])
    if str(code.SCHEME) == 'batch':
        # west inclination: equal to the UNMODIFIED reference on every shot
        corr, _, _ = orc.decode_batch(0, FLAGS['west'], arr['syn_bits'], want_growth=False)
        np.testing.assert_array_equal(orc.logical_batch(arr['err_bits'], corr), arr['west_logical_unmodified'])


SYNDROME_7F = {(0, 0), (0, 1), (2, 0), (0, 4), (1, 4), (3, 4), (4, 4), (4, 3)}
# tests/decoders/uf/conftest.py:5-44 (reference fixture ``validated_static_erasure``)
STATIC_ERASURE_7F = {
    ((0, 0), (0, 1)), ((0, 3), (0, 4)), ((0, 4), (0, 5)), ((0, 4), (1, 4)), ((1, 0), (2, 0)),
    ((1, 4), (2, 4)), ((2, -1), (2, 0)), ((2, 0), (2, 1)), ((2, 0), (3, 0)), ((2, 3), (2, 4)),
    ((2, 3), (3, 3)), ((2, 4), (2, 5)), ((2, 4), (3, 4)), ((2, 5), (3, 5)), ((3, 2), (3, 3)),
    ((3, 2), (4, 2)), ((3, 3), (3, 4)), ((3, 3), (4, 3)), ((3, 4), (3, 5)), ((3, 4), (4, 4)),
    ((3, 5), (3, 6)), ((3, 5), (4, 5)), ((4, 1), (4, 2)), ((4, 2), (4, 3)), ((4, 2), (5, 2)),
    ((4, 3), (4, 4)), ((4, 3), (5, 3)), ((4, 4), (4, 5)), ((4, 4), (5, 4)), ((4, 5), (4, 6)),
    ((4, 5), (5, 5)), ((5, 2), (5, 3)), ((5, 3), (5, 4)), ((5, 3), (6, 3)), ((5, 4), (5, 5)),
    ((5, 4), (6, 4))}


def test_uf_known_answers_from_reference_tests():
    """Surface(7, code capacity), ``syndrome7F`` (tests/decoders/conftest.py:18-28):
    |erasure| = 38 (tests/decoders/uf/UF/test_UF_unit.py:76-85), the exact static
    erasure (tests/decoders/uf/conftest.py:5-44; the fixture lists 36 of the 38
    edges, so it is checked as a subset), a syndrome-valid correction inside the
    erasure (test_UF_unit.py:174-178), 29 full edges when dynamic
    (test_UF_unit.py:87-96).  The reference's |correction| == 6
    (test_UF_integration.py:21-25) is a property of CPython's set order in
    ``_make_tree``; the canonical edge-id order peels the same erasure with 8
    edges (checked against oracle/ref_canonical.py when generating the goldens)."""
    import localuf_b200 as L
    code = L.Surface(7, 'code capacity')
    g = code.FLAT
    orc = binding.Oracle.from_code(code)
    syn = g.syndrome_to_bits(SYNDROME_7F)[None, :]
    corr, growth, _ = orc.decode_batch(0, 0, syn)
    erasure = {g.edges[k] for k in np.flatnonzero(growth[0] == 2)}
    assert len(erasure) == 38
    assert STATIC_ERASURE_7F <= erasure
    correction = g.bits_to_edges(corr[0])
    assert len(correction) == 8 and correction <= erasure
    assert code.get_syndrome(correction) == SYNDROME_7F
    _, growth, _ = orc.decode_batch(0, 1, syn)
    assert int(np.count_nonzero(growth[0] == 2)) == 29


def test_uf_single_defect_exact_corrections():
    """One defect next to a boundary has a unique correction (cf. the reference's
    tests/decoders/_inclination/*_integration.py)."""
    import localuf_b200 as L
    code = L.Surface(5, 'code capacity')
    g = code.FLAT
    orc = binding.Oracle.from_code(code)
    for flags in (0, 2):
        corr, _, _ = orc.decode_batch(0, flags, g.syndrome_to_bits({(2, 0)})[None, :])
        assert g.bits_to_edges(corr[0]) == {((2, -1), (2, 0))}
        corr, _, _ = orc.decode_batch(0, flags, g.syndrome_to_bits({(2, 3)})[None, :])
        assert g.bits_to_edges(corr[0]) == {((2, 3), (2, 4))}
        corr, _, _ = orc.decode_batch(0, flags, g.syndrome_to_bits(set())[None, :])
        assert g.bits_to_edges(corr[0]) == set()


def test_mc_batch_threads_and_rates():
    """oracle_mc_batch: p = 0 never fails; the d=5 repetition code at p = 0.05
    fails at the exact majority-vote rate (UF corrects every weight <= 2 error
    and fails on every weight >= 3 error) within 5 sigma."""
    import localuf_b200 as L
    code = L.Repetition(5, 'code capacity')
    orc = binding.Oracle.from_code(code)
    assert orc.mc_batch(0, 0, [0.0], 1, 2000, nthreads=2) == 0
    n, p = 200000, 0.05
    m = orc.mc_batch(0, 0, [p], 7, n, nthreads=4)
    exact = sum(comb(5, k) * p**k * (1 - p)**(5 - k) for k in range(3, 6))
    sigma = (exact * (1 - exact) / n) ** 0.5
    assert abs(m / n - exact) < 5 * sigma
