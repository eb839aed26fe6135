"""The C oracle of the decoder confidence scores (oracle/luf_oracle_dcs.c) against the values the unmodified
reference computed (tests/golden/dcs_*.npz: UF; dcsf_*.npz: Snowflake, per decoding cycle), and the host
tables (weights, meta-boundary sides, growth membership) against the reference's."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

TOL = dict(rtol=1e-12, atol=1e-12)        # floating point: sums in a different order than networkx / Python's sum


@pytest.mark.parametrize('name', golden_names('dcs_'))
def test_oracle_dcs_matches_reference_uf(name):
    from localuf_b200 import _dcs
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    for level, tag in ((None, 'none'), (meta['p'], 'p')):
        t = _dcs.flat_tables(code, level)
        if level is not None:
            np.testing.assert_allclose(t.weight, arr['weight_p'], rtol=1e-15, atol=0)
        for variant in ('static', 'dynamic'):
            out = binding.dcs_batch(t, arr[f'{variant}_growth'])
            np.testing.assert_allclose(out['swim'], arr[f'{variant}_swim_{tag}'], **TOL)
            np.testing.assert_allclose(out['uef'], arr[f'{variant}_uef_{tag}'], **TOL)
            np.testing.assert_allclose(1 - out['clustered'] / code.NODE_COUNT, arr[f'{variant}_unf'], **TOL)
            if level is None:                      # unit weights: half-integers, exact
                np.testing.assert_array_equal(out['swim'], arr[f'{variant}_swim_none'])


@pytest.mark.parametrize('name', golden_names('dcsf_'))
def test_oracle_dcs_matches_reference_snowflake(name):
    from localuf_b200 import _dcs
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    h, ta = code.SCHEME.WINDOW_HEIGHT, code.TIME_AXIS
    for level, tag in ((None, 'none'), (meta['p'], 'p')):
        t = _dcs.flat_tables(code, level)
        t.in_growth = np.array([all(v[ta] < h for v in e) for e in code.EDGES], np.uint8)
        np.testing.assert_array_equal(t.in_growth, arr['in_growth'])
        if level is not None:
            np.testing.assert_allclose(t.weight, arr['weight_p'], rtol=1e-15, atol=0)
        out = binding.dcs_batch(t, arr['growth'], want=('swim', 'uef'))
        np.testing.assert_allclose(out['swim'], arr[f'swim_{tag}'], **TOL)
        np.testing.assert_allclose(out['uef'], arr[f'uef_{tag}'], **TOL)


@pytest.mark.parametrize('name', golden_names('dcsf_'))
def test_window_tables_are_the_flat_tables_reindexed(name):
    """The layered edge ids of the resident window carry the same endpoints, weights and membership."""
    from localuf_b200 import _dcs, _window
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    w = _window.build(code)
    t = _dcs.window_tables(code, w, meta['p'])
    ELp = 32 * ((w.EL + 31) // 32)
    assert t.n_edges == w.h * ELp and t.growth_words * 16 == t.n_edges
    for k in range(len(code.EDGES)):
        b, l = divmod(int(w.layered_of_code[k]), w.EL)
        j = b * ELp + l
        assert (t.edge_u[j], t.edge_v[j]) == (code.FLAT.edge_u[k], code.FLAT.edge_v[k])
        assert t.weight[j] == arr['weight_p'][k] or abs(t.weight[j] - arr['weight_p'][k]) < 1e-15
        assert t.in_growth[j] == arr['in_growth'][k]
    holes = np.ones(t.n_edges, bool)
    holes[[(int(x) // w.EL) * ELp + int(x) % w.EL for x in w.layered_of_code]] = False
    assert not t.weight[holes].any() and not t.in_growth[holes].any()


def test_pack_growth_round_trip():
    from localuf_b200 import _dcs, _engine
    rng = np.random.default_rng(5)
    codes = rng.integers(0, 4, size=(7, 53), dtype=np.uint8)
    np.testing.assert_array_equal(_engine.unpack_growth(_dcs.pack_growth(codes), 53), codes)
