"""GPU parity tests of the decoder confidence scores (csrc/luf_dcs.cuh through luf_dcs_run; the per-cycle
export of the stream kernels through luf_stream_*_metrics) against the values the unmodified reference
computed (tests/golden/dcs_*.npz, dcsf_*.npz), the C oracle on GPU samples, and the known answers of the
reference's own tests (tests/decoders/test_BaseUF.py:75-112, tests/decoders/uf/UF/test_UF_integration.py:5-17)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu
TOL = dict(rtol=1e-12, atol=1e-12)      # float64 sums in a different order (block reduction / Bellman-Ford vs networkx)


@pytest.mark.parametrize('name', golden_names('dcs_'))
def test_dcs_kernel_matches_reference_uf(name):
    from localuf_b200 import _dcs
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    for level, tag in ((None, 'none'), (meta['p'], 'p')):
        eng = _dcs.DcsEngine(_dcs.flat_tables(code, level))
        for variant in ('static', 'dynamic'):
            swim, uef, cl = eng.run_host(_dcs.pack_growth(arr[f'{variant}_growth']))
            np.testing.assert_allclose(swim, arr[f'{variant}_swim_{tag}'], **TOL)
            np.testing.assert_allclose(uef, arr[f'{variant}_uef_{tag}'], **TOL)
            np.testing.assert_allclose(1 - cl / code.NODE_COUNT, arr[f'{variant}_unf'], **TOL)
            if level is None:
                np.testing.assert_array_equal(swim, arr[f'{variant}_swim_none'])       # half-integers: exact


@pytest.mark.parametrize('name', golden_names('dcs_'))
def test_uf_score_methods_after_decode(name):
    """decode(syndrome) then swim_distance / unclustered_* on the decoder object, as a user of the reference would."""
    import localuf_b200 as L
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    flat = code.FLAT
    dec = L.decoders.UF(code)
    for s in range(0, meta['shots'], max(1, meta['shots'] // 8)):
        dec.reset()
        dec.decode(flat.bits_to_syndrome(arr['syn_bits'][s]))
        assert dec.swim_distance() == arr['static_swim_none'][s]
        assert abs(dec.swim_distance(noise_level=meta['p']) - arr['static_swim_p'][s]) < 1e-12
        assert abs(dec.unclustered_edge_fraction() - arr['static_uef_none'][s]) < 1e-12
        assert abs(dec.unclustered_edge_fraction(meta['p']) - arr['static_uef_p'][s]) < 1e-12
        assert abs(dec.unclustered_node_fraction() - arr['static_unf'][s]) < 1e-12


def _set_growth(dec, edge, code):
    dec._growth_codes[dec.CODE.FLAT.edge_index[edge]] = code


def test_reference_known_answers_swim_distance():
    import localuf_b200 as L
    uf = L.decoders.UF(L.Surface(5, 'code capacity'))
    assert uf.swim_distance() == 5                                       # test_no_clusters_5F
    for i in range(5):
        for max_j in range(5):
            for code_, want in ((2, 5 - max_j), (1, 5 - max_j / 2)):     # test_percolated_5F / test_half_percolated_5F
                uf.reset()
                for j in range(-1, max_j - 1):
                    _set_growth(uf, ((i, j), (i, j + 1)), code_)
                assert uf.swim_distance() == want
    uf = L.decoders.UF(L.Surface(3, 'phenomenological'))
    assert uf.swim_distance() == 3                                       # test_no_clusters_3T
    for i in range(3):
        for max_j in range(3):
            for t in range(3):
                for code_, want in ((2, 3 - max_j), (1, 3 - max_j / 2)):
                    uf.reset()
                    for j in range(-1, max_j - 1):
                        _set_growth(uf, ((i, j, t), (i, j + 1, t)), code_)
                    assert uf.swim_distance() == want
    dec = L.decoders.UF(L.Surface(6, 'code capacity'))                   # test_Meister2024_example
    dec.decode({(2, 1), (3, 3)})
    assert dec.swim_distance() == 1
    uf7 = L.decoders.UF(L.Surface(7, 'code capacity'))                   # test_syndrome7F
    uf7.validate({(0, 0), (0, 1), (2, 0), (0, 4), (1, 4), (3, 4), (4, 4), (4, 3)})
    assert uf7.swim_distance() == 2
    with pytest.raises(ValueError):
        uf7.complementary_gap()                                          # test_UF_unit.py:223-226


@pytest.mark.parametrize('spec,p,n', [
    (dict(code='Surface', d=7, noise='phenomenological'), 0.02, 600),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.01, 600),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.01, 300),
    (dict(code='Surface', d=17, noise='phenomenological'), 0.02, 40),
])
def test_sampled_scores_vs_oracle(spec, p, n):
    """sample -> decode -> check -> scores without leaving the device, against the oracle on the same shots."""
    import localuf_b200 as L
    from localuf_b200 import _dcs, _engine
    code = make_code(spec)
    dec = L.decoders.UF(code)
    out = dec.sample_confidence_scores(p, n, noise_level_for_priors=p, seed=11, chunk=256)
    eng = dec.engine
    err, syn = eng.sample_host(p, 11, 0, n)
    corr, growth2b, _ = eng.decode_batch_host(0, 0, syn, want_growth=True)
    logical, _ = eng.check_host(err, corr)
    np.testing.assert_array_equal(out['logical'], logical)
    ref = binding.dcs_batch(_dcs.flat_tables(code, p), _engine.unpack_growth(growth2b, code.N_EDGES))
    np.testing.assert_allclose(out['swim_distance'], ref['swim'], **TOL)
    np.testing.assert_allclose(out['unclustered_edge_fraction'], ref['uef'], **TOL)
    np.testing.assert_allclose(out['unclustered_node_fraction'], 1 - ref['clustered'] / code.NODE_COUNT, **TOL)
    assert out['swim_distance'].min() >= 0 and out['swim_distance'].max() <= code.D * _dcs.flat_tables(code, p).weight.max()


def flags_of(kw):
    return (1 if kw.get('merger') == 'slow' else 0) | (2 if kw.get('schedule') == '1:1' else 0)


@pytest.mark.parametrize('tile', ['warp', 'cta'])
@pytest.mark.parametrize('name', golden_names('dcsf_'))
def test_snowflake_cycle_metrics_match_reference(name, tile, monkeypatch):
    """luf_stream_decode_metrics + luf_dcs_run on the reference's error stream: every per-cycle score of
    Snowflake.decode(metrics=...) (snowflake/main.py:266-297)."""
    import localuf_b200 as L
    from localuf_b200 import _dcs
    monkeypatch.setenv('LUF_SF_TILE', tile)
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    dec = L.decoders.Snowflake(code, **meta['decoder_kwargs'])
    t = dec._tables
    fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
    rt, _, _, growth, mins = dec.engine.decode_metrics(flags_of(meta['decoder_kwargs']), fresh[None])
    np.testing.assert_array_equal(rt[0, :, 0], arr['runtime_all'])
    np.testing.assert_array_equal(mins[0, :, 0], arr['min_defect_height'])
    np.testing.assert_array_equal(mins[0, :, 1], arr['min_active_layer'])
    for level, tag in ((None, 'none'), (meta['p'], 'p')):
        swim, uef, _ = dec._scores(growth, level)
        np.testing.assert_allclose(swim, arr[f'swim_{tag}'], **TOL)
        np.testing.assert_allclose(uef, arr[f'uef_{tag}'], **TOL)
        if level is None:
            np.testing.assert_array_equal(swim, arr['swim_none'])


def test_snowflake_decode_with_metrics_api():
    """Snowflake.decode(syndrome, metrics=...) cycle by cycle, and the score methods on the decoder."""
    import localuf_b200 as L
    meta, arr = load_golden('dcsf_sf3_cl')
    code = make_code(meta['spec'])
    dec = L.decoders.Snowflake(code)
    t = dec._tables
    fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
    metrics = ('throughput', 'swim_distance', 'min_defect_height', 'unclustered_edge_fraction', 'min_active_layer')
    ncyc = 12
    # Frugal._load (_schemes.py:545-564): a fresh edge flips its top-layer detectors now and its
    # future-boundary detectors in the next cycle
    top = [v for v in code.NODES if v[-1] == t.h - 1 and not code.is_boundary(v)]
    carried = set()
    for c in range(ncyc):
        now, nxt = set(carried), set()
        for l in range(t.EL):
            if (fresh[c][l >> 5] >> (l & 31)) & 1:
                for q, dt in zip(t.edge_q[l], t.edge_dt[l]):
                    if q >= t.NLb:
                        (nxt if dt else now).symmetric_difference_update({int(q)})
        carried = nxt
        syndrome = {v for v in top if t.position_of(v) in now}
        rt = dec.decode(syndrome, metrics=metrics, noise_level_for_priors=meta['p'], time_only='all')
        assert rt == arr['runtime_all'][c]
    h = dec.confidence_score_history
    np.testing.assert_allclose(h['swim_distance'], arr['swim_p'][:ncyc], **TOL)
    np.testing.assert_allclose(h['unclustered_edge_fraction'], arr['uef_p'][:ncyc], **TOL)
    assert h['min_defect_height'] == arr['min_defect_height'][:ncyc].tolist()
    assert h['throughput'] == [1 / x for x in arr['runtime_all'][:ncyc]]
    assert sum(dec.min_active_layers.values()) == ncyc
    assert dec.min_active_layer() == arr['min_active_layer'][ncyc - 1]
    assert dec.swim_distance() == arr['swim_none'][ncyc - 1]
    g = dec.growth
    assert [int(g[e]) for e in code.EDGES if e in g] == arr['growth'][ncyc - 1][arr['in_growth'].astype(bool)].tolist()


def test_frugal_run_with_metrics():
    """Frugal.run(decoder, p, n, metrics=...): one entry per timed cycle and metric; the swim distance is
    that of the exported growth (checked through the oracle), the histogram counts every timed cycle."""
    import localuf_b200 as L
    from localuf_b200 import _dcs, _engine
    code = L.Surface(5, 'circuit-level', scheme='frugal')
    dec = L.decoders.Snowflake(code)
    n, shots, p = 3, 7, 0.012
    metrics = ('throughput', 'swim_distance', 'min_defect_height', 'unclustered_edge_fraction', 'min_active_layer')
    m, _ = code.SCHEME.run(dec, p, n, shots=shots, seed=5, time_only='all', metrics=metrics)
    m_plain, _ = code.SCHEME.run(L.decoders.Snowflake(code), p, n, shots=shots, seed=5, time_only='all')
    assert m == m_plain
    h = dec.confidence_score_history
    cycles = shots * n * code.D
    assert all(len(h[k]) == cycles for k in metrics if k != 'min_active_layer')
    assert sum(dec.min_active_layers.values()) == cycles
    counts, steps, growth, cycle = dec.engine.mc_metrics(0, p, 5, 0, shots, n)
    ref = binding.dcs_batch(_dcs.window_tables(code, dec._tables, p),
                            _engine.unpack_growth(growth.cpu().numpy().view(np.uint32), dec._tables.h * 32 * dec.engine.W),
                            want=('swim', 'uef'))
    np.testing.assert_allclose(h['swim_distance'], ref['swim'], **TOL)
    np.testing.assert_allclose(h['unclustered_edge_fraction'], ref['uef'], **TOL)
    assert h['throughput'] == [1 / r for r in cycle.reshape(-1, 4)[:, 2]]
    assert h['min_defect_height'] == cycle.reshape(-1, 4)[:, 0].tolist()
    np.testing.assert_array_equal(cycle[:, :, 2].reshape(shots, n, code.D).sum(axis=2), steps[:, :, 0])
    assert list(code.SCHEME.step_counts) == steps[:, :, 0].reshape(-1).tolist()
    with pytest.raises(ValueError):
        code.SCHEME.run(dec, p, 1, metrics=('foo',))
