"""GPU parity tests of Snowflake + the frugal scheme through the C ABI."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu


def flags_of(kw):
    return (1 if kw.get('merger') == 'slow' else 0) | (2 if kw.get('schedule') == '1:1' else 0)


@pytest.mark.parametrize('name', golden_names('sf_'))
def test_sf_matches_reference_goldens(name):
    from localuf_b200 import _engine
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    eng = _engine.StreamEngine(code, device=0)
    t = eng.tables
    fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
    rt, floor, m = eng.decode_host(flags_of(meta['decoder_kwargs']), fresh[None])
    np.testing.assert_array_equal(rt[0, :, 0], arr['runtime_all'])
    np.testing.assert_array_equal(rt[0, :, 1], arr['runtime_unrooting'])
    np.testing.assert_array_equal(floor[0], binding.code_bits_to_local(t, arr['floor_bits'], t.h - 1))
    np.testing.assert_array_equal(m[0], arr['m_cycle'])
    assert int(m.sum()) == meta['m'] == meta['unmodified_m']


@pytest.mark.parametrize('spec,p,flags,nstreams,ncycles', [
    (dict(code='Repetition', d=7, noise='phenomenological', kwargs=dict(scheme='frugal')), 0.08, 0, 300, 60),
    (dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), 0.03, 0, 200, 40),
    (dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.012, 0, 200, 40),
    (dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.012, 3, 100, 40),
    (dict(code='Surface', d=9, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.006, 0, 60, 40),
    (dict(code='Surface', d=13, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.004, 0, 20, 45),
])
def test_sf_streams_vs_oracle_on_random_errors(spec, p, flags, nstreams, ncycles):
    """Many windows in parallel on seeded host errors (cfg4 size included): every
    stream's per-cycle runtimes, floor corrections and logical counts equal the oracle's."""
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = _engine.StreamEngine(code, device=0)
    t = eng.tables
    orc = binding.StreamOracle(t, merger='slow' if flags & 1 else 'fast', schedule='1:1' if flags & 2 else '2:1')
    rng = np.random.default_rng(2026)
    probs = np.array(code.NOISE.class_probabilities(p))[t.edge_class]
    flips = rng.random((nstreams, ncycles, t.EL)) < probs
    flips[:, -2 * t.h:, :] = False                                    # cleanse phase
    pad = np.zeros((nstreams, ncycles, 32 * eng.W), np.uint8)
    pad[:, :, :t.EL] = flips
    fresh = np.packbits(pad, axis=2, bitorder='little').view(np.uint32)
    rt, floor, m = eng.decode_host(flags, fresh)
    for k in range(nstreams):
        o_all, o_unr, o_floor, o_m = orc.stream(fresh[k])
        np.testing.assert_array_equal(rt[k, :, 0], o_all)
        np.testing.assert_array_equal(rt[k, :, 1], o_unr)
        np.testing.assert_array_equal(floor[k], o_floor)
        np.testing.assert_array_equal(m[k], o_m)


def test_sf_mc_rates_and_api():
    import localuf_b200 as L
    code = L.Surface(5, 'circuit-level', scheme='frugal')
    dec = L.decoders.Snowflake(code)
    t = dec._tables
    orc = binding.StreamOracle(t)
    p, n = 0.012, 2
    counts, steps = dec.engine.mc_host(0, p, 9, 0, 4000, n, want_steps=True)
    total = t.h + n * 5 + 5 + 2 * t.h
    assert counts[1] == 4000 * total and steps.shape == (4000, n, 2)
    m_cpu, cyc_cpu = orc.mc(code.NOISE.class_probabilities(p), 10, 2000, n, nthreads=len(os.sched_getaffinity(0)))
    assert cyc_cpu == 2000 * total
    r_gpu, r_cpu = counts[0] / 4000, m_cpu / 2000
    sigma = ((r_gpu + r_cpu) / 2 * (1 / 4000 + 1 / 2000)) ** 0.5 + 1e-9
    assert abs(r_gpu - r_cpu) < 5 * sigma, (r_gpu, r_cpu)
    # no noise: nothing to correct, idle runtimes only
    assert dec.engine.mc_host(0, 0.0, 1, 0, 10, 1)[0][0] == 0
    # drop-in API
    m, slenderness = code.SCHEME.run(dec, p, 3)
    assert slenderness == t.h / 5 + 3 + 1 and len(code.SCHEME.step_counts) == 3 and m >= 0
    m, slenderness = code.SCHEME.run(dec, p, 2, shots=50, time_only='all')
    assert slenderness == 50 * (t.h / 5 + 2 + 1) and len(code.SCHEME.step_counts) == 100
    assert min(code.SCHEME.step_counts) >= 4 * 5
    with pytest.raises(ValueError):
        code.SCHEME.run(dec, p, 0)
    with pytest.raises(ValueError):
        L.decoders.Snowflake(L.Surface(3, 'phenomenological'))
    with pytest.raises(NotImplementedError):
        L.decoders.Snowflake(code, unrooter='simple')


@pytest.fixture
def block_tile(monkeypatch):
    """One thread block per window (luf::sf_cta), whatever the window size."""
    monkeypatch.setenv('LUF_SF_TILE', 'cta')


@pytest.mark.parametrize('name', ['sf_sf5_cl', 'sf_sf5_cl_slow11', 'sf_rep5_phen', 'sf_sf7_cl', 'sf_sf5_phen_b2'])
def test_sf_block_tile_matches_reference_goldens(name, block_tile):
    test_sf_matches_reference_goldens(name)


@pytest.mark.parametrize('spec,p,flags,nstreams,ncycles', [
    (dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.03, 3, 100, 40),
    (dict(code='Surface', d=9, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.008, 0, 60, 40),
    (dict(code='Surface', d=13, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.007, 0, 20, 45),
])
def test_sf_block_tile_streams_vs_oracle(spec, p, flags, nstreams, ncycles, block_tile):
    test_sf_streams_vs_oracle_on_random_errors(spec, p, flags, nstreams, ncycles)


def test_sf_large_window_uses_blocks_and_matches_oracle():
    """d = 17 (53 KB of state per window): the block tile is the default; streams of seeded errors
    near threshold equal the oracle's."""
    test_sf_streams_vs_oracle_on_random_errors(
        dict(code='Surface', d=17, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.006, 0, 6, 40)


def test_sf_mc_counts_do_not_depend_on_the_tile(monkeypatch):
    """The sampler keeps 32 lanes whatever the tile: same seed, same memory experiments, same counts."""
    import localuf_b200 as L
    code = L.Surface(9, 'circuit-level', scheme='frugal')
    dec = L.decoders.Snowflake(code)
    out = {}
    for tile in ('warp', 'cta'):
        monkeypatch.setenv('LUF_SF_TILE', tile)
        out[tile] = dec.engine.mc_host(0, 0.007, 31, 5, 600, 2, want_steps=True)
    assert out['warp'][0] == out['cta'][0] and out['warp'][0][0] > 0
    np.testing.assert_array_equal(out['warp'][1], out['cta'][1])
