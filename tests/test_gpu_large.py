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


def test_global_batch_window_beyond_16bit_layout():
    """Global batch scheme (_schemes.py:166-203) with h = d * n = 1650 layers: 33000 detectors, 100630 edges.
    Sampled, decoded (wide instance) and counted (k_count_strings_wide) on the device; corrections equal the
    oracle's, per-window counts equal the host mirror of Global.get_logical_error on the same windows."""
    import localuf_b200 as L
    from localuf_b200 import _engine
    n = 330
    code = L.Surface(5, 'phenomenological', scheme='global batch', window_height=5 * n)
    assert code.FLAT.n_detectors > 32766
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    p, shots = 0.01, 12
    err, syn = eng.sample_host(p, 5, 0, shots)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    corr, _, steps = eng.decode_batch_host(_engine.DEC_UF, 0, syn)
    ocorr, _, osteps = orc.decode_batch(0, 0, syn, want_growth=False)
    np.testing.assert_array_equal(corr, ocorr)
    np.testing.assert_array_equal(steps[:, 0], osteps[:, 0])
    total, per = eng.mc_global(_engine.DEC_UF, 0, p, 5, 0, shots, chunk=8, want_counts=True)
    np.testing.assert_array_equal(per, [code.SCHEME.count_leftover_bits(l) for l in err ^ corr])
    assert total == int(per.sum())
    m, slender = code.SCHEME.run(L.decoders.UF(code), p, n, shots=4, seed=5)
    assert slender == 4 * n and m >= 0


@pytest.mark.parametrize('d,noise,p,nstreams,ncycles', [
    (23, 'circuit-level', 0.004, 6, 5),          # 135608 edges, 25898 nodes: beyond the 16-bit layout
    (25, 'phenomenological', 0.02, 6, 5),        # 33100 nodes: beyond the 16-bit layout
    (17, 'phenomenological', 0.02, 12, 6),       # fits the 16-bit indices, but a stream's state exceeds an SM's shared memory
])
def test_forward_windows_beyond_16bit_layout(d, noise, p, nstreams, ncycles):
    """The forward scheme (_schemes.py:282-450) around the wide UF instance (luf_fw_wide.cuh): per-cycle round
    counts, commit leftovers and logical counts equal the oracle's on the same fresh errors; Forward.run runs."""
    import localuf_b200 as L
    from localuf_b200 import _engine
    code = L.Surface(d, noise, scheme='forward')
    eng = _engine.Engine(code, device=0)
    ft = eng.attach_forward()
    orc = binding.ForwardOracle(code, ft)
    g = code.FLAT
    rng = np.random.default_rng(11)
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
    for flags in (0, 2):
        steps, commit, m = eng.forward_decode_host(_engine.DEC_UF, flags, init_b, fresh_b)
        for k in range(nstreams):
            o_steps, o_commit, o_m = orc.stream(0, flags, init_b[k], fresh_b[k])
            np.testing.assert_array_equal(steps[k], o_steps)
            np.testing.assert_array_equal(commit[k], o_commit)
            np.testing.assert_array_equal(m[k], o_m)
    m, slender = code.SCHEME.run(L.decoders.UF(code), p / 4, 2, shots=4)
    assert m >= 0 and slender > 0
