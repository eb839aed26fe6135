"""Snowflake / frugal kernel phases (localuf_b200/csrc/luf_sf.cuh compiled for
the host) vs. goldens exported from the unmodified reference.  CPU-only."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import binding  # noqa: E402
import emu_binding  # noqa: E402


def flags_of(kw):
    return (1 if kw.get('merger') == 'slow' else 0) | (2 if kw.get('schedule') == '1:1' else 0)


@pytest.mark.parametrize('name', golden_names('sf_'))
def test_emu_sf_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.EmuStream(code)
    t = emu.tables
    fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
    rt, floor, m = emu.stream(flags_of(meta['decoder_kwargs']), fresh)
    np.testing.assert_array_equal(rt[:, 0], arr['runtime_all'], err_msg=f'{name} runtime (all)')
    np.testing.assert_array_equal(rt[:, 1], arr['runtime_unrooting'], err_msg=f'{name} runtime (unrooting)')
    np.testing.assert_array_equal(floor, binding.code_bits_to_local(t, arr['floor_bits'], t.h - 1))
    np.testing.assert_array_equal(m, arr['m_cycle'])


def test_emu_sf_mc_runs_and_is_quiet_without_noise():
    code = make_code(dict(code='Surface', d=3, noise='circuit-level', kwargs=dict(scheme='frugal')))
    emu = emu_binding.EmuStream(code)
    h, d, n = emu.tables.h, 3, 2
    counts, steps = emu.mc(0, 0.0, 5, 0, 3, n)
    total = h + n * d + d + 2 * h
    assert counts[0] == 0 and counts[1] == 3 * total
    assert counts[2] == 3 * n * d * 4 and counts[3] == 0       # idle cycle: runtime 'all' = 2 loops x (1 + 1)
    assert (steps[:, :, 0] == 4 * d).all()
    counts, _ = emu.mc(0, 0.02, 5, 0, 40, n)
    assert counts[1] == 40 * total and counts[2] > 40 * n * d * 4


def test_emu_sf_bitmap_sweep_with_tiny_awake_list():
    """A build whose awake-node list holds 4 entries: the sweeps fall back to the bitmap."""
    emu_binding.use_tiny_lists(True)
    try:
        for name in ('sf_sf5_cl', 'sf_sf5_phen', 'sf_rep5_phen'):
            meta, arr = load_golden(name)
            emu = emu_binding.EmuStream(make_code(meta['spec']))
            t = emu.tables
            fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
            rt, floor, m = emu.stream(flags_of(meta['decoder_kwargs']), fresh)
            np.testing.assert_array_equal(rt[:, 0], arr['runtime_all'])
            np.testing.assert_array_equal(floor, binding.code_bits_to_local(t, arr['floor_bits'], t.h - 1))
            np.testing.assert_array_equal(m, arr['m_cycle'])
    finally:
        emu_binding.use_tiny_lists(False)


@pytest.mark.parametrize('spec,p,flags,nstreams,ncycles', [
    (dict(code='Repetition', d=7, noise='phenomenological', kwargs=dict(scheme='frugal')), 0.08, 0, 40, 60),
    (dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='frugal')), 0.03, 0, 30, 40),
    (dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.012, 0, 30, 40),
    (dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.03, 3, 20, 40),
    (dict(code='Surface', d=7, noise='circuit-level', kwargs=dict(scheme='frugal')), 0.008, 0, 10, 45),
])
def test_emu_sf_streams_vs_oracle_on_random_errors(spec, p, flags, nstreams, ncycles):
    """Seeded host errors, dense enough that clusters merge, unroot and leave through the floor
    while others are still growing: per-cycle runtimes, floor corrections and logical counts of
    the kernel phases equal the C oracle's on every stream."""
    code = make_code(spec)
    emu = emu_binding.EmuStream(code)
    t = emu.tables
    orc = binding.StreamOracle(t, merger='slow' if flags & 1 else 'fast', schedule='1:1' if flags & 2 else '2:1')
    rng = np.random.default_rng(77)
    probs = np.array(code.NOISE.class_probabilities(p))[t.edge_class]
    flips = rng.random((nstreams, ncycles, t.EL)) < probs
    flips[:, -2 * t.h:, :] = False
    pad = np.zeros((nstreams, ncycles, 32 * emu.W), np.uint8)
    pad[:, :, :t.EL] = flips
    fresh = np.packbits(pad, axis=2, bitorder='little').view(np.uint32)
    for k in range(nstreams):
        rt, floor, m = emu.stream(flags, fresh[k])
        o_all, o_unr, o_floor, o_m = orc.stream(fresh[k])
        np.testing.assert_array_equal(rt[:, 0], o_all, err_msg=f'stream {k}')
        np.testing.assert_array_equal(rt[:, 1], o_unr, err_msg=f'stream {k}')
        np.testing.assert_array_equal(floor, o_floor, err_msg=f'stream {k}')
        np.testing.assert_array_equal(m, o_m, err_msg=f'stream {k}')


@pytest.mark.parametrize('nl', [64, 256])
@pytest.mark.parametrize('name', ['sf_sf5_cl', 'sf_sf5_cl_slow11', 'sf_rep5_phen', 'sf_sf7_cl', 'sf_sf5_phen_b2'])
def test_emu_sf_block_tile_matches_reference(name, nl):
    """The block-per-window instance of the drivers (luf::sf_cta; tiles of 64 and 256 lanes)."""
    meta, arr = load_golden(name)
    emu = emu_binding.EmuStream(make_code(meta['spec']))
    t = emu.tables
    fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
    rt, floor, m = emu.stream(flags_of(meta['decoder_kwargs']), fresh, nl=nl)
    np.testing.assert_array_equal(rt[:, 0], arr['runtime_all'])
    np.testing.assert_array_equal(rt[:, 1], arr['runtime_unrooting'])
    np.testing.assert_array_equal(floor, binding.code_bits_to_local(t, arr['floor_bits'], t.h - 1))
    np.testing.assert_array_equal(m, arr['m_cycle'])


def test_emu_sf_mc_does_not_depend_on_the_tile():
    """The sampler keeps 32 lanes whatever the tile, so a seeded memory experiment gives the same
    counts and step sums on warp and block tiles."""
    code = make_code(dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='frugal')))
    emu = emu_binding.EmuStream(code)
    warp = emu.mc(0, 0.012, 9, 100, 12, 2)
    for nl in (64, 256):
        block = emu.mc(0, 0.012, 9, 100, 12, 2, nl=nl)
        assert block[0] == warp[0]
        np.testing.assert_array_equal(block[1], warp[1])
    assert warp[0][0] > 0


def layered_growth_to_code_order(t, growth_words, n_code_edges):
    """[n, h * 2 * W(EL)] exported growth words -> [n, E] codes in code-edge order."""
    from localuf_b200 import _engine
    ELp = 32 * ((t.EL + 31) // 32)
    codes = _engine.unpack_growth(growth_words, t.h * ELp)
    idx = np.array([(int(x) // t.EL) * ELp + int(x) % t.EL for x in t.layered_of_code[:n_code_edges]])
    return codes[:, idx]


@pytest.mark.parametrize('nl', [32, 128])
@pytest.mark.parametrize('name', golden_names('dcsf_'))
def test_emu_sf_cycle_metrics_match_reference(name, nl):
    """Per-cycle export of the stream drivers: growth of the window, min_defect_height and min_active_layer
    (snowflake/main.py:266-297) vs the unmodified reference; the scores themselves through the oracle."""
    from localuf_b200 import _dcs
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.EmuStream(code)
    t = emu.tables
    fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
    rt, _, _, growth, mins = emu.stream_metrics(flags_of(meta['decoder_kwargs']), fresh, nl=nl)
    np.testing.assert_array_equal(rt[:, 0], arr['runtime_all'])
    np.testing.assert_array_equal(layered_growth_to_code_order(t, growth, meta['E']), arr['growth'])
    np.testing.assert_array_equal(mins[:, 0], arr['min_defect_height'])
    np.testing.assert_array_equal(mins[:, 1], arr['min_active_layer'])
    wt = _dcs.window_tables(code, t, meta['p'])
    from localuf_b200 import _engine
    out = binding.dcs_batch(wt, _engine.unpack_growth(growth, wt.n_edges), want=('swim', 'uef'))
    np.testing.assert_allclose(out['swim'], arr['swim_p'], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(out['uef'], arr['uef_p'], rtol=1e-12, atol=1e-12)
