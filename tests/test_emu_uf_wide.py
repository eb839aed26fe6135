"""The wide instance of the UF kernel phases (luf::uf_wide, localuf_b200/csrc/luf_uf_wide.cuh: 64-bit node
words and 32-bit indices for graphs beyond the 16-bit layout) compiled for the host, vs. the reference goldens
and the C oracle.  Every golden graph fits the wide layout too, so the same cases pin it.  CPU-only."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import binding  # noqa: E402
import emu_binding  # noqa: E402
from localuf_b200._engine import unpack_growth  # noqa: E402

FLAGS = {'default': 0, 'dynamic': 1, 'west': 2}


@pytest.mark.parametrize('name', golden_names('uf_'))
@pytest.mark.parametrize('nl', [32, 256])
def test_emu_wide_matches_reference(name, nl):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    for variant, flags in FLAGS.items():
        corr, growth2b, steps = emu.decode_uf_wide(flags, arr['syn_bits'], nl=nl)
        np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])


@pytest.mark.parametrize('name', golden_names('nuf_') + golden_names('nbuf_') + golden_names('buf_'))
def test_emu_wide_equals_narrow_on_node_and_bucket_decoders(name):
    """NodeUF / NodeBUF / BUF flags: the wide instance equals the 16-bit one (itself pinned by the goldens)."""
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    base = {'nuf_': 4, 'nbuf_': 12, 'buf_': 8}[name[:name.index('_') + 1]]
    for flags in ([base, base | 1, base | 2] if base == 4 else [base]):
        a = emu.decode_uf(flags, arr['syn_bits'])
        b = emu.decode_uf_wide(flags, arr['syn_bits'], nl=64)
        for x, y in zip(a, b):
            np.testing.assert_array_equal(x, y)


@pytest.mark.parametrize('spec,p', [
    (dict(code='Repetition', d=5, noise='code capacity'), 0.05),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02),
    (dict(code='Surface', d=9, noise='phenomenological'), 0.025),
])
@pytest.mark.parametrize('flags', [0, 2])
def test_emu_wide_fused_vs_oracle(spec, p, flags):
    """Sample -> decode -> logical bit: same random stream as the 16-bit kernels, corrections and per-shot bit
    equal to the oracle's; the parity-only peel gives the same bit as the full one."""
    code = make_code(spec)
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    n = 200
    fails, err, syn, corr, logical = emu.mc_uf_wide(flags, p, seed=99, shot0=5, nshots=n)
    _, err16, syn16, _ = emu.mc_uf(flags, p, seed=99, shot0=5, nshots=n)
    np.testing.assert_array_equal(err, err16)
    np.testing.assert_array_equal(syn, syn16)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    ocorr, _, _ = orc.decode_batch(0, flags, syn, want_growth=False)
    np.testing.assert_array_equal(ocorr, corr)
    ological = orc.logical_batch(err, ocorr)
    np.testing.assert_array_equal(logical, ological)
    assert fails == int(ological.sum())
    fails2, _, _, _, logical2 = emu.mc_uf_wide(flags, p, seed=99, shot0=5, nshots=n, parity_only=True)
    np.testing.assert_array_equal(logical2, ological)
