"""Subset (fixed-weight) sampling, SURVEY.md section 8 row f3:
``Noise.force_error`` / ``subset_probabilities`` (localuf/noise/main.py:38-66,116-134,
noise/forcers.py), ``Batch.sim_cycles_given_weight`` (_schemes.py:157-163) and
``Decoder.subset_sample`` (_base_classes.py:682-712).  Goldens: oracle/gen_golden_subset.py."""
import itertools
import os
import random
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))


@pytest.mark.parametrize('name', golden_names('ss_'))
def test_host_subset_api_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    noise = code.NOISE
    assert len(noise.ALL_WEIGHTS) == meta['n_all_weights']
    if 'subset_prob' in arr:
        df = noise.subset_probabilities(meta['p'])
        assert list(df.columns) == ['subset prob', 'survival prob']
        index = np.array([list(w) if isinstance(w, tuple) else [w] for w in df.index], np.int64)
        np.testing.assert_array_equal(index, arr['index'])
        np.testing.assert_allclose(df['subset prob'].to_numpy(), arr['subset_prob'], rtol=1e-13, atol=0)
        np.testing.assert_allclose(df['survival prob'].to_numpy(), arr['survival_prob'], rtol=1e-13, atol=1e-300)
    # same global random stream, same call sequence -> the very same forced errors
    random.seed(meta['seed'])
    k = 0
    for w in meta['weights']:
        w = tuple(w) if isinstance(w, list) else w
        for _ in range(6):
            np.testing.assert_array_equal(code.FLAT.edges_to_bits(noise.force_error(w)), arr['forced'][k])
            k += 1
    with pytest.raises(ValueError):
        noise.force_error(len(code.EDGES) + 1 if isinstance(meta['weights'][0], int)
                          else tuple(10 ** 6 for _ in meta['weights'][0]))


def test_emu_fixed_weight_sampler_logic():
    """ph_force compiled for the host: exactly w distinct fresh edges, uniform, syndrome consistent."""
    import binding
    import emu_binding
    code = make_code(dict(code='Surface', d=3, noise='phenomenological'))
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    w, n = 5, 6000
    err, syn = emu.sample_weight(w, 3, 0, n)
    bits = np.unpackbits(err.view(np.uint8), axis=1, bitorder='little')[:, :code.N_EDGES]
    assert (bits.sum(axis=1) == w).all()
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    q = w / code.N_EDGES
    assert np.all(np.abs(bits.mean(axis=0) - q) < 6 * (q * (1 - q) / n) ** 0.5)
    full, _ = emu.sample_weight(code.N_EDGES, 3, 0, 3)
    assert (np.unpackbits(full.view(np.uint8), axis=1, bitorder='little').sum(axis=1) == code.N_EDGES).all()


def exact_failure_fraction(code, orc, weight):
    """Failure fraction of the canonical UF over ALL errors of the given weight (oracle, tiny codes)."""
    E = code.N_EDGES
    errs = []
    for combo in itertools.combinations(range(E), weight):
        errs.append(code.FLAT.edges_to_bits({code.EDGES[k] for k in combo}))
    err = np.stack(errs)
    corr, _, _ = orc.decode_batch(0, 0, orc.syndrome(err), want_growth=False)
    return float(orc.logical_batch(err, corr).mean())


@pytest.mark.gpu
def test_device_fixed_weight_sampler():
    import binding
    from localuf_b200 import _engine
    for spec, w in ((dict(code='Surface', d=3, noise='code capacity'), 3),
                    (dict(code='Surface', d=5, noise='phenomenological'), 7),
                    (dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='forward')), 4)):
        code = make_code(spec)
        eng = _engine.Engine(code, device=0)
        orc = binding.Oracle.from_code(code)
        n = 30000
        err, syn = eng.sample_weight_host(w, 7, 100, n)
        bits = np.unpackbits(err.view(np.uint8), axis=1, bitorder='little')[:, :code.N_EDGES]
        assert (bits.sum(axis=1) == w).all()                        # exactly w distinct edges
        np.testing.assert_array_equal(orc.syndrome(err), syn)
        fresh = np.array([code.FLAT.edge_index[e] for e in code.NOISE.FRESH_EDGES])
        assert not np.delete(bits, fresh, axis=1).any()             # only fresh edges
        q = w / len(fresh)
        freq = bits[:, fresh].mean(axis=0)
        assert np.all(np.abs(freq - q) < 6 * (q * (1 - q) / n) ** 0.5)   # every edge equally likely
        # pairs of edges too (no position bias from the take-t-or-j rule)
        a, b = fresh[0], fresh[-1]
        both = (bits[:, a] & bits[:, b]).mean()
        q2 = q * (w - 1) / (len(fresh) - 1)
        assert abs(both - q2) < 6 * (q2 / n) ** 0.5 + 1e-4
        # counter-based: any split of the shot range gives the same samples
        e2, _ = eng.sample_weight_host(w, 7, 1100, 50)
        np.testing.assert_array_equal(e2, err[1000:1050])
    eng = _engine.Engine(make_code(dict(code='Surface', d=3, noise='code capacity')), device=0)
    e0, s0 = eng.sample_weight_host(0, 1, 0, 10)
    assert not e0.any() and not s0.any()
    eall, _ = eng.sample_weight_host(13, 1, 0, 10)
    assert (np.unpackbits(eall.view(np.uint8), axis=1, bitorder='little').sum(axis=1) == 13).all()
    with pytest.raises(ValueError):
        eng.sample_weight_host(14, 1, 0, 10)


@pytest.mark.gpu
@pytest.mark.parametrize('decoder', ['UF', 'Macar'])
def test_sim_cycles_given_weight_vs_exact_enumeration(decoder):
    import binding
    import localuf_b200 as L
    code = L.Surface(3, 'code capacity')
    orc = binding.Oracle.from_code(code)
    dec = getattr(L.decoders, decoder)(code)
    n = 40000
    for w in (1, 2, 3):
        m, shots = code.SCHEME.sim_cycles_given_weight(dec, w, n, seed=11 + w)
        assert shots == n
        if decoder == 'UF':
            f = exact_failure_fraction(code, orc, w)
            assert abs(m / n - f) < 5 * (max(f * (1 - f), 1e-4) / n) ** 0.5, (w, m / n, f)
        elif w == 1:
            assert m == 0                                            # d = 3 corrects every single error
    with pytest.raises(ValueError):
        code.SCHEME.sim_cycles_given_weight(dec, 14, 10)


@pytest.mark.gpu
def test_subset_sample_dataframe_and_estimate():
    import localuf_b200 as L
    code = L.Surface(5, 'code capacity')
    uf = L.decoders.UF(code)
    p, n = 0.06, 20000
    random.seed(3)
    df = uf.subset_sample(p, n)
    assert list(df.columns) == ['subset prob', 'survival prob', 'm', 'n']
    probs = code.NOISE.subset_probabilities(p)
    assert list(df.index) == list(probs.index[:len(df)])              # most probable subsets first
    rows = df.dropna()
    assert (rows['n'] == n).all() and np.isnan(df.loc[0, 'm'])
    estimate = float((rows['subset prob'] * rows['m'] / rows['n']).sum())
    m, _ = code.SCHEME.run(uf, p, 2_000_000, seed=9)
    direct = m / 2_000_000
    assert abs(estimate - direct) < 0.15 * direct                       # cut-off subsets carry < tol * mean
    # circuit-level weights are tuples; errors come from the reference's host sampler
    cl = L.Surface(3, 'circuit-level')
    macar = L.decoders.Macar(cl)
    random.seed(4)
    m, shots = cl.SCHEME.sim_cycles_given_weight(macar, (1, 1), 300)
    assert shots == 300 and 0 <= m <= 300
    # the sweep driver (localuf/sim/accuracy/monte_carlo.py:204-233)
    random.seed(5)
    df = L.sim.accuracy.subset_sample([3, 5], 0.05, 2000, L.Surface, L.decoders.UF, 'code capacity')
    assert df.index.names[0] == 'd' and set(df.index.get_level_values(0)) == {3, 5}
    assert list(df.columns) == ['subset prob', 'survival prob', 'm', 'n']
