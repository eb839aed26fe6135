"""GPU parity tests of the forward scheme (UF / Macar) through the C ABI."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402
from localuf_b200 import _window  # noqa: E402

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize('name', golden_names('fw_'))
def test_forward_matches_reference_goldens(name):
    from localuf_b200 import _engine
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    eng = _engine.Engine(code, device=0)
    for key, dec in (('uf', 0), ('macar', 1)):
        steps, commit, m = eng.forward_decode_host(dec, 0, arr[f'{key}_init'][None], arr[f'{key}_fresh'][None])
        np.testing.assert_array_equal(steps[0], arr[f'{key}_steps'])
        np.testing.assert_array_equal(commit[0], arr[f'{key}_commit'])
        np.testing.assert_array_equal(m[0], arr[f'{key}_m'])
    assert meta['m']['macar'] == meta['unmodified_macar_m']


@pytest.mark.parametrize('spec,p,nstreams,ncycles', [
    (dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='forward')), 0.03, 200, 12),
    (dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='forward')), 0.012, 100, 10),
    (dict(code='Surface', d=9, noise='phenomenological', kwargs=dict(scheme='forward')), 0.025, 60, 8),
    (dict(code='Surface', d=7, noise='phenomenological',
          kwargs=dict(scheme='forward', commit_height=3, buffer_height=7)), 0.03, 60, 12),
])
def test_forward_streams_vs_oracle(spec, p, nstreams, ncycles):
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    ft = eng.attach_forward()
    orc = binding.ForwardOracle(code, ft)
    g = code.FLAT
    rng = np.random.default_rng(7)
    probs = np.array(code.NOISE.class_probabilities(p))

    def bits(mask):
        pad = np.zeros(mask.shape[:-1] + (32 * eng.WE,), np.uint8)
        pad[..., :g.n_edges] = mask
        return np.packbits(pad, axis=-1, bitorder='little').view(np.uint32)

    fresh_p = np.where(g.edge_class != 255, probs[np.minimum(g.edge_class, len(probs) - 1)], 0.0)
    buf_p = np.where(ft.buffer_class != 255, probs[np.minimum(ft.buffer_class, len(probs) - 1)], 0.0)
    fresh = rng.random((nstreams, ncycles, g.n_edges)) < fresh_p
    fresh[:, -ft.n_cleanse:, :] = False
    init = rng.random((nstreams, g.n_edges)) < buf_p
    fresh_b, init_b = bits(fresh), bits(init)
    for dec in (0, 1):
        steps, commit, m = eng.forward_decode_host(dec, 0, init_b, fresh_b)
        for k in range(nstreams):
            o_steps, o_commit, o_m = orc.stream(dec, 0, init_b[k], fresh_b[k])
            np.testing.assert_array_equal(steps[k], o_steps)
            np.testing.assert_array_equal(commit[k], o_commit)
            np.testing.assert_array_equal(m[k], o_m)


def test_forward_mc_and_api():
    import localuf_b200 as L
    code = L.Surface(5, 'phenomenological', scheme='forward')
    uf, macar = L.decoders.UF(code), L.decoders.Macar(code)
    assert code.SCHEME.run(uf, 0.0, 3) == (0, (5 + 4 * 5) / 5)               # tests/_schemes/Forward: run(uf, 0, n) == (0, n+2)
    assert code.SCHEME.step_counts == [None, None, None]
    m, sl = code.SCHEME.run(macar, 0.0, 4)
    assert (m, sl) == (0, 6.0) and code.SCHEME.step_counts == [(4, 2)] * 4      # idle Macar decode = (4, 2)
    m_uf, sl = code.SCHEME.run(uf, 0.03, 10, shots=400)
    m_mc, _ = code.SCHEME.run(macar, 0.03, 10, shots=400)
    assert sl == 400 * 12.0 and len(code.SCHEME.step_counts) == 4000
    assert m_uf > 0 and m_mc > 0 and 0.3 < m_uf / m_mc < 3.0                      # same order of magnitude
    df = L.sim.runtime.forward([3, 5], [0.01, 0.02], 6, 'phenomenological')
    assert df.shape == (6, 8) and list(df.columns.names) == ['stage', 'd', 'p']
    with pytest.raises(ValueError):
        code.SCHEME.run(uf, 0.01, 0)
