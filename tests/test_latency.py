"""``Frugal.sample_latency`` / ``sim.latency.frugal`` / ``sim.runtime.frugal`` (reference
``_schemes.py:651-698``, ``sim/latency.py:13-72``, ``sim/runtime.py:139-199``).

Goldens (``oracle/gen_golden_lat.py``): the errors the unmodified reference drew in every cycle of
``sample_latency`` and the latency it returned for the three ``time_only`` modes.  Replayed through
the given-errors path with ``LUF_SF_LATENCY_TAIL`` (the noiseless tail in which the decoder stops
growing once no defect is left): CPU on the host build of the kernel drivers, GPU through the C ABI."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import binding  # noqa: E402
import localuf_b200 as L  # noqa: E402
from localuf_b200 import _engine  # noqa: E402


def flags_of(kw):
    return (1 if kw.get('merger') == 'slow' else 0) | (2 if kw.get('schedule') == '1:1' else 0)


def latencies(rt, h, loops):
    """Per-cycle runtimes ('all', 'unrooting') of one experiment -> latency in the three modes.
    A cycle of 1 timestep is a drop without growth or merging (defects_possible = False)."""
    tail = rt[-h:]
    full = tail[:, 0] > 1
    return int(tail[:, 0].sum()), int((tail[:, 0] - 2 * loops)[full].sum()), int(tail[:, 1].sum())


def check_case(name, decode):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    from localuf_b200 import _window
    t = _window.build(code)
    flags = flags_of(meta['decoder_kwargs']) | _engine.SF_LATENCY_TAIL
    loops = 1 if meta['decoder_kwargs'].get('schedule') == '1:1' else 2
    assert arr['fresh_bits'].shape[1] == 2 * t.h + 1
    got = []
    for sample in arr['fresh_bits']:
        rt = decode(code, flags, binding.code_bits_to_local(t, sample, 0))
        got.append(latencies(rt, t.h, loops))
    got = np.array(got)
    np.testing.assert_array_equal(got[:, 0], arr['latency_all'], err_msg=f'{name} all')
    np.testing.assert_array_equal(got[:, 1], arr['latency_merging'], err_msg=f'{name} merging')
    np.testing.assert_array_equal(got[:, 2], arr['latency_unrooting'], err_msg=f'{name} unrooting')


def on_emulation(nl):
    import emu_binding
    cache = {}

    def decode(code, flags, fresh):
        emu = cache.setdefault(id(code), emu_binding.EmuStream(code))
        return emu.stream(flags, fresh, nl=nl)[0]
    return decode


def on_device(code, flags, fresh):
    return _engine.StreamEngine(code).decode_host(flags, fresh[None], want_floor=False)[0][0]


@pytest.mark.parametrize('name', golden_names('lat_'))
def test_emu_latency_matches_reference(name):
    check_case(name, on_emulation(32))


@pytest.mark.parametrize('name', ['lat_sf5_cl', 'lat_rep7_phen_11'])
def test_emu_latency_block_tile(name):
    check_case(name, on_emulation(64))


def test_emu_fused_latency_agrees_with_replay_statistics():
    """The fused kernel driver (device sampler) against the golden latencies: same distribution
    (means within 5 sigma), identical on warp and block tiles, zero noise gives the idle value."""
    import emu_binding
    meta, arr = load_golden('lat_sf3_cl')
    code = make_code(meta['spec'])
    emu = emu_binding.EmuStream(code)
    h = emu.tables.h
    lat = emu.latency(0, meta['p'], 5, 0, 400)
    np.testing.assert_array_equal(lat, emu.latency(0, meta['p'], 5, 0, 400, nl=64))
    for col, key in enumerate(('latency_all', 'latency_merging', 'latency_unrooting')):
        ref = arr[key].astype(float)
        sigma = (ref.var() / len(ref) + lat[:, col].var() / len(lat)) ** 0.5 + 1e-9
        assert abs(ref.mean() - lat[:, col].mean()) < 5 * sigma + 0.5, key
    quiet = emu.latency(0, 0.0, 1, 0, 3)
    # one full idle cycle (2 loops x (grow + 1 merging step)), then h - 1 drops
    assert (quiet == [4 + (h - 1), 0, 0]).all()


@pytest.mark.gpu
@pytest.mark.parametrize('name', golden_names('lat_'))
def test_gpu_latency_matches_reference(name):
    check_case(name, on_device)


@pytest.mark.gpu
def test_gpu_sample_latency_api(monkeypatch):
    code = L.Surface(5, 'circuit-level', scheme='frugal')
    dec = L.decoders.Snowflake(code)
    h = code.SCHEME.WINDOW_HEIGHT
    assert code.SCHEME.sample_latency(dec, 0.0, time_only='all') == 4 + (h - 1)
    assert code.SCHEME.sample_latency(dec, 0.0) == 0
    many = code.SCHEME.sample_latency(dec, 0.012, time_only='all', shots=2000, seed=3)
    assert len(many) == 2000 and min(many) >= h + 3 and max(many) > 4 + (h - 1)
    merging = code.SCHEME.sample_latency(dec, 0.012, shots=2000, seed=3)
    unrooting = code.SCHEME.sample_latency(dec, 0.012, time_only='unrooting', shots=2000, seed=3)
    assert all(a >= m >= u >= 0 for a, m, u in zip(many, merging, unrooting))
    # the same seed gives the same experiments on warp and block tiles
    monkeypatch.setenv('LUF_SF_TILE', 'cta')
    assert code.SCHEME.sample_latency(dec, 0.012, time_only='all', shots=2000, seed=3) == many
    monkeypatch.delenv('LUF_SF_TILE')
    # against the emulation-free statistic of the golden replay: same mean within 5 sigma
    meta, arr = load_golden('lat_sf5_cl')
    ref = arr['latency_merging'].astype(float)
    got = np.array(merging, float)
    sigma = (ref.var() / len(ref) + got.var() / len(got)) ** 0.5
    assert abs(ref.mean() - got.mean()) < 5 * sigma + 0.5
    with pytest.raises(ValueError):
        code.SCHEME.sample_latency(dec, 0.01, time_only='nothing')
    # drivers: DataFrame shapes of the reference
    df = L.sim.latency.frugal([3, 5], [0.005, 0.01], 50, L.Surface, 'circuit-level')
    assert df.shape == (50, 4) and list(df.columns.names) == ['d', 'p']
    df = L.sim.runtime.frugal([3, 5], [0.005, 0.01], 7, L.Surface, 'circuit-level', time_only='all')
    assert df.shape == (7, 4) and list(df.columns.names) == ['d', 'p'] and (df.values >= 4 * 3).all()
