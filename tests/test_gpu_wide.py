"""The wide instance of the UF kernels (localuf_b200/csrc/luf_uf_wide.cuh) on the GPU, forced onto graphs that
also fit the 16-bit layout (LUF_FORCE_WIDE=1) so that every reference golden pins it: growth, rounds and
corrections of the staged decode, the sampler's random stream, the fused kernel's per-shot bit, the string
counter.  State slab in shared and in global memory (LUF_WIDE_STATE)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu
FLAGS = {'default': 0, 'dynamic': 1, 'west': 2}


@pytest.fixture
def wide(monkeypatch):
    monkeypatch.setenv('LUF_FORCE_WIDE', '1')
    return monkeypatch


@pytest.mark.parametrize('name', golden_names('uf_'))
@pytest.mark.parametrize('state', ['smem', 'global'])
def test_wide_decode_matches_reference(name, state, wide):
    from localuf_b200 import _engine
    wide.setenv('LUF_WIDE_STATE', state)
    meta, arr = load_golden(name)
    eng = _engine.Engine(make_code(meta['spec']), device=0)
    for variant, flags in FLAGS.items():
        corr, growth2b, steps = eng.decode_batch_host(_engine.DEC_UF, flags, arr['syn_bits'], want_growth=True)
        np.testing.assert_array_equal(_engine.unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])


@pytest.mark.parametrize('name', golden_names('nuf_') + golden_names('nbuf_') + golden_names('buf_'))
def test_wide_node_and_bucket_decoders_equal_16bit_kernels(name, monkeypatch):
    from localuf_b200 import _engine
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    base = {'nuf_': 4, 'nbuf_': 12, 'buf_': 8}[name[:name.index('_') + 1]]
    narrow = _engine.Engine(code, device=0)
    monkeypatch.setenv('LUF_FORCE_WIDE', '1')
    monkeypatch.setenv('LUF_WIDE_STATE', 'global')
    wide_eng = _engine.Engine(code, device=0)
    for flags in ([base, base | 1, base | 2] if base == 4 else [base]):
        a = narrow.decode_batch_host(_engine.DEC_UF, flags, arr['syn_bits'], want_growth=True)
        b = wide_eng.decode_batch_host(_engine.DEC_UF, flags, arr['syn_bits'], want_growth=True)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize('spec,p', [
    (dict(code='Repetition', d=7, noise='phenomenological'), 0.06),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.026),
    (dict(code='Surface', d=17, noise='phenomenological'), 0.02),
])
@pytest.mark.parametrize('state', ['smem', 'global'])
def test_wide_sampler_and_fused_kernel_vs_oracle(spec, p, state, monkeypatch):
    from localuf_b200 import _engine
    code = make_code(spec)
    narrow = _engine.Engine(code, device=0)
    monkeypatch.setenv('LUF_FORCE_WIDE', '1')
    monkeypatch.setenv('LUF_WIDE_STATE', state)
    monkeypatch.setenv('LUF_FAST', '0')              # the wide fused kernel itself, not the compact fast launch
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    n, seed, shot0 = 400, 5, 123
    err, syn = eng.sample_host(p, seed, shot0, n)
    err16, syn16 = narrow.sample_host(p, seed, shot0, n)
    np.testing.assert_array_equal(err, err16)
    np.testing.assert_array_equal(syn, syn16)
    for flags in (0, 1, 2, 4, 12):
        ocorr, _, _ = orc.decode_batch(0, flags, syn, want_growth=False)
        want = orc.logical_batch(err, ocorr)
        logical, fails, shots = eng.mc_logical_host(_engine.DEC_UF, flags, p, seed, shot0, n)
        np.testing.assert_array_equal(logical, want)
        assert (fails, shots) == (int(want.sum()), n)
    # fixed-weight sampler: same picks as the 16-bit kernel
    e1, s1 = eng.sample_weight_host(6, seed, shot0, 100)
    e2, s2 = narrow.sample_weight_host(6, seed, shot0, 100)
    np.testing.assert_array_equal(e1, e2)
    np.testing.assert_array_equal(s1, s2)


@pytest.mark.parametrize('spec,p', [
    (dict(code='Surface', d=5, noise='phenomenological'), 0.04),
    (dict(code='Surface', d=3, noise='circuit-level'), 0.02),
])
def test_wide_string_counter_equals_16bit_kernel(spec, p, monkeypatch):
    from localuf_b200 import _engine
    code = make_code(spec)
    narrow = _engine.Engine(code, device=0)
    monkeypatch.setenv('LUF_FORCE_WIDE', '1')
    eng = _engine.Engine(code, device=0)
    err, syn = narrow.sample_host(p, 3, 0, 500)
    corr, _, _ = narrow.decode_batch_host(_engine.DEC_UF, 0, syn)
    np.testing.assert_array_equal(eng.count_strings_host(err, corr), narrow.count_strings_host(err, corr))


@pytest.mark.parametrize('name', golden_names('fw_'))
@pytest.mark.parametrize('state', ['smem', 'global'])
def test_wide_forward_matches_reference_goldens(name, state, wide):
    """The forward scheme around the wide UF instance (luf_fw_wide.cuh) on the reference's forward goldens."""
    from localuf_b200 import _engine
    wide.setenv('LUF_WIDE_STATE', state)
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    eng = _engine.Engine(code, device=0)
    steps, commit, m = eng.forward_decode_host(0, 0, arr['uf_init'][None], arr['uf_fresh'][None])
    np.testing.assert_array_equal(steps[0], arr['uf_steps'])
    np.testing.assert_array_equal(commit[0], arr['uf_commit'])
    np.testing.assert_array_equal(m[0], arr['uf_m'])


def test_wide_forward_mc_equals_16bit_kernels(monkeypatch):
    """Forward.run with device-sampled noise: same random stream, same counts as the 16-bit forward kernels."""
    from localuf_b200 import _engine
    code = make_code(dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='forward')))
    narrow = _engine.Engine(code, device=0)
    a, sa = narrow.forward_mc_host(_engine.DEC_UF, 0, 0.01, 9, 0, 64, 6, want_steps=True)
    monkeypatch.setenv('LUF_FORCE_WIDE', '1')
    eng = _engine.Engine(code, device=0)
    b, sb = eng.forward_mc_host(_engine.DEC_UF, 0, 0.01, 9, 0, 64, 6, want_steps=True)
    assert a == b
    np.testing.assert_array_equal(sa, sb)
