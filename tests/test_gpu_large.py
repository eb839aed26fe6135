"""Graphs beyond the 16-bit layout of the canonical kernels: Surface(23 | 25,
'circuit-level') -- 65869 / 85297 edges (localuf/codes.py:155-251).  The compact path (luf_uf_fast.cuh) moves
the edge / other-end split of its adjacency entries and decodes every shot itself under the west inclination,
whose logical outcome equals the unmodified reference's on every shot."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('d,p,n', [(23, 0.004, 300), (25, 0.004, 200), (25, 0.007, 100)])
def test_wide_graph_west_logical_bit_per_shot(d, p, n):
    import localuf_b200 as L
    from localuf_b200 import _engine
    code = L.Surface(d, 'circuit-level')
    assert code.N_EDGES > 65533
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    seed, shot0 = 31, 1000
    err, syn = eng.sample_host(p, seed, shot0, n)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    corr, _, _ = orc.decode_batch(0, 2, syn)
    want = orc.logical_batch(err, corr)
    logical, fails, shots = eng.mc_logical_host(_engine.DEC_UF, _engine.UF_WEST, p, seed, shot0, n)
    np.testing.assert_array_equal(logical, want)
    assert (fails, shots) == (int(want.sum()), n)
    # K4 on the wide edge arrays: error against the oracle's correction
    bits, k4 = eng.check_host(err, corr)
    np.testing.assert_array_equal(bits, want)
    assert k4 == fails
    # the class API: UF(inclination='west') through Batch.run
    m, slender = code.SCHEME.run(L.decoders.UF(code, inclination='west'), p, 2000, seed=3)
    assert slender == 2000 and 0 <= m <= 2000
    # Macar / Actis keep their 16-bit cellular-automaton layout
    with pytest.raises(NotImplementedError):
        eng.decode_batch_host(_engine.DEC_MACAR, 0, syn)


@pytest.mark.parametrize('d,p,n', [(23, 0.004, 60), (25, 0.005, 40)])
@pytest.mark.parametrize('state', ['global', 'smem'])
def test_wide_graph_canonical_decode_vs_oracle(d, p, n, state, monkeypatch):
    """The wide instance of the canonical kernels (luf_uf_wide.cuh: 64-bit node words, the state slab in global
    memory or, where it fits, in shared memory): growth, rounds, corrections and the logical bit of the staged
    decode equal the oracle's under the default, dynamic and west settings; the fused kernel's per-shot bit too
    (default inclination: compact fast launch + wide launch over the HARD shots; dynamic: the wide kernel alone)."""
    import localuf_b200 as L
    from localuf_b200 import _engine
    monkeypatch.setenv('LUF_WIDE_STATE', state)
    code = L.Surface(d, 'circuit-level')
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    seed, shot0 = 77, 0
    err, syn = eng.sample_host(p, seed, shot0, n)
    for flags in (0, 1, 2):
        corr, growth2b, steps = eng.decode_batch_host(_engine.DEC_UF, flags, syn, want_growth=True)
        ocorr, ogrowth, osteps = orc.decode_batch(0, flags, syn)
        np.testing.assert_array_equal(_engine.unpack_growth(growth2b, code.N_EDGES), ogrowth)
        np.testing.assert_array_equal(steps[:, 0], osteps[:, 0])
        np.testing.assert_array_equal(corr, ocorr)
        want = orc.logical_batch(err, ocorr)
        logical, fails, shots = eng.mc_logical_host(_engine.DEC_UF, flags, p, seed, shot0, n)
        np.testing.assert_array_equal(logical, want)
        assert (fails, shots) == (int(want.sum()), n)
    m, slender = code.SCHEME.run(L.decoders.UF(code), p, 500, seed=3)
    assert slender == 500 and 0 <= m <= 500
