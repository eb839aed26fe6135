"""GPU parity tests of Macar / Actis through the C ABI: bit-exact timestep
counts, erasure and corrections vs. reference goldens and vs. the C oracle on
GPU-sampled errors."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu
DECODERS = {'macar': (1, 0), 'actis': (2, 0), 'actis_unopt': (2, 4)}


def engine_of(code):
    from localuf_b200 import _engine
    return _engine.Engine(code, device=0)


@pytest.mark.parametrize('name', golden_names('ca_'))
def test_ca_matches_reference_goldens(name):
    from localuf_b200 import _engine
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    eng = engine_of(code)
    for key in meta['decoders']:
        dec, flags = DECODERS[key]
        corr, growth2b, steps = eng.decode_batch_host(dec, flags, arr['syn_bits'], want_growth=True)
        np.testing.assert_array_equal(steps, arr[f'{key}_steps'])
        np.testing.assert_array_equal(_engine.unpack_growth(growth2b, meta['E']), arr[f'{key}_growth'])
        np.testing.assert_array_equal(corr, arr[f'{key}_corr'])
        if key == 'macar' and str(code.SCHEME) == 'batch':
            logical, _ = eng.check_host(arr['err_bits'], corr)
            np.testing.assert_array_equal(logical, arr['macar_logical'])


@pytest.mark.parametrize('spec,p,n_macar,n_actis', [
    (dict(code='Surface', d=5, noise='code capacity'), 0.08, 4000, 1000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.06, 3000, 300),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.03, 2000, 300),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.01, 1000, 200),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.01, 600, 24),
])
def test_ca_pipeline_vs_oracle_on_gpu_samples(spec, p, n_macar, n_actis):
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = engine_of(code)
    orc = binding.Oracle.from_code(code)
    for key, n in (('macar', n_macar), ('actis', n_actis), ('actis_unopt', max(n_actis // 2, 8))):
        dec, flags = DECODERS[key]
        err, syn = eng.sample_host(p, 77, 5000, n)
        corr, growth2b, steps = eng.decode_batch_host(dec, flags, syn, want_growth=True)
        ocorr, ogrowth, osteps = orc.decode_batch(dec, flags, syn)
        np.testing.assert_array_equal(steps, osteps)
        np.testing.assert_array_equal(_engine.unpack_growth(growth2b, code.N_EDGES), ogrowth)
        np.testing.assert_array_equal(corr, ocorr)
        if key == 'macar':
            np.testing.assert_array_equal(orc.syndrome(corr), syn)
            logical, fails = eng.check_host(err, corr)
            np.testing.assert_array_equal(logical, orc.logical_batch(err, ocorr))
            fused, shots = eng.mc_run_host(dec, flags, p, 77, 5000, n)
            assert (fused, shots) == (fails, n)
            assert eng.last_step_sums == (int(steps[:, 0].sum()), int(steps[:, 1].sum()))
        else:
            assert not corr.any()
            eng.mc_run_host(dec, flags, p, 77, 5000, n)
            assert eng.last_step_sums == (int(steps[:, 0].sum()), int(steps[:, 1].sum()))


def test_ca_python_api():
    import localuf_b200 as L
    code = L.Surface(5, 'code capacity')
    macar = L.decoders.Macar(code)
    assert macar.decode(set()) == (4, 2)                     # SURVEY appendix A3 known answers
    macar.reset()
    assert macar.decode({(2, 1)}) == (30, 6)
    assert macar.correction == {((2, -1), (2, 0)), ((2, 0), (2, 1))}
    macar.reset()
    assert macar.decode({(2, 1), (2, 2)}) == (6, 3)
    actis = L.decoders.Actis(code)
    assert actis.decode(set()) == (35, 22) and actis.correction == set()
    assert L.decoders.Actis(L.Surface(5, 'phenomenological')).decode(set()) == (47, 30)
    assert L.decoders.Actis(L.Surface(5, 'circuit-level')).decode(set()) == (23, 14)
    m, n = code.SCHEME.run(macar, 0.05, 20000)
    assert n == 20000 and 0 < m < 20000
    df = L.sim.runtime.batch([3, 5], [0.02, 0.05], 50, 'code capacity', validate_only=False)
    assert df.shape == (50, 8) and list(df.columns.names) == ['stage', 'd', 'p']
    with pytest.raises(ValueError):
        L.decoders.Actis(L.Surface(3, 'phenomenological', scheme='forward'))
    with pytest.raises(ValueError):
        L.decoders.Macar(L.Surface(3, 'phenomenological', scheme='frugal'))


@pytest.fixture(params=[128, 512])
def block_tile(request, monkeypatch):
    """One thread block per shot (luf::ca_cta), whatever the graph size."""
    monkeypatch.setenv('LUF_CA_TILE', 'cta')
    monkeypatch.setenv('LUF_CA_CTA_THREADS', str(request.param))


@pytest.mark.parametrize('name', golden_names('ca_'))
def test_ca_block_tile_matches_reference_goldens(name, block_tile):
    test_ca_matches_reference_goldens(name)


@pytest.mark.parametrize('spec,p,n_macar,n_actis', [
    (dict(code='Surface', d=5, noise='circuit-level'), 0.01, 1000, 200),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.01, 600, 48),
])
def test_ca_block_tile_pipeline_vs_oracle(spec, p, n_macar, n_actis, block_tile):
    test_ca_pipeline_vs_oracle_on_gpu_samples(spec, p, n_macar, n_actis)


def test_ca_mc_counts_do_not_depend_on_the_tile(monkeypatch):
    """Same seed, same shots, same failure count and timestep sums on either tile."""
    from localuf_b200 import _engine
    code = make_code(dict(code='Surface', d=7, noise='phenomenological'))
    eng = engine_of(code)
    out = {}
    for tile in ('warp', 'cta'):
        monkeypatch.setenv('LUF_CA_TILE', tile)
        for dec in (1, 2):
            out[tile, dec] = (eng.mc_run_host(dec, 0, 0.02, 5, 0, 3000), eng.last_step_sums)
    assert out['warp', 1] == out['cta', 1] and out['warp', 1][0][0] > 0
    assert out['warp', 2] == out['cta', 2]
