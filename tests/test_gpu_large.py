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
    # the class API: UF(inclination='west') through Batch.run; the default inclination needs the canonical kernels
    m, slender = code.SCHEME.run(L.decoders.UF(code, inclination='west'), p, 2000, seed=3)
    assert slender == 2000 and 0 <= m <= 2000
    with pytest.raises(NotImplementedError):
        code.SCHEME.run(L.decoders.UF(code), p, 100)
    with pytest.raises(NotImplementedError):
        eng.decode_batch_host(_engine.DEC_UF, 2, syn)
