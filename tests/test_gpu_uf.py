"""GPU parity tests of the UF path, all through the C ABI (libluf_b200.so).

Bit-exact against (a) golden vectors exported from the reference and (b) the C
oracle re-deriving everything from the very errors the GPU sampled; failure
rates against exact values / the oracle's own Monte Carlo within binomial
confidence intervals.
"""
import os
import random
import sys
from math import comb

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu

FLAGS = {'default': 0, 'dynamic': 1, 'west': 2}


def engine_of(code):
    from localuf_b200 import _engine
    return _engine.Engine(code, device=0)


@pytest.mark.parametrize('name', golden_names('uf_'))
def test_decode_matches_reference_goldens(name):
    from localuf_b200 import _engine
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    eng = engine_of(code)
    for variant, flags in FLAGS.items():
        corr, growth2b, steps = eng.decode_batch_host(_engine.DEC_UF, flags, arr['syn_bits'], want_growth=True)
        np.testing.assert_array_equal(_engine.unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])
        if str(code.SCHEME) == 'batch':
            logical, fails = eng.check_host(arr['err_bits'], corr)
            np.testing.assert_array_equal(logical, arr[f'{variant}_logical'])
            assert fails == int(arr[f'{variant}_logical'].sum())
    if str(code.SCHEME) == 'batch':      # west inclination == UNMODIFIED reference, every shot
        corr, _, _ = eng.decode_batch_host(_engine.DEC_UF, FLAGS['west'], arr['syn_bits'])
        logical, _ = eng.check_host(arr['err_bits'], corr)
        np.testing.assert_array_equal(logical, arr['west_logical_unmodified'])


CASES = [
    (dict(code='Repetition', d=5, noise='code capacity'), 0.05, 20000),
    (dict(code='Repetition', d=9, noise='phenomenological'), 0.04, 5000),
    (dict(code='Surface', d=3, noise='code capacity'), 0.12, 20000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.05, 20000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.10, 20000),
    (dict(code='Surface', d=7, noise='code capacity'), 0.30, 5000),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.03, 10000),
    (dict(code='Surface', d=7, noise='phenomenological', kwargs=dict(window_height=3)), 0.03, 5000),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.01, 20000),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.03, 5000),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.01, 5000),
    (dict(code='Surface', d=7, noise='circuit-level', kwargs=dict(parametrization='standard')), 0.005, 3000),
    (dict(code='Surface', d=13, noise='phenomenological'), 0.02, 2000),
    # large graphs: one thread block per shot (luf::uf_cta) is chosen automatically
    (dict(code='Surface', d=15, noise='phenomenological'), 0.02, 600),
    (dict(code='Surface', d=19, noise='phenomenological'), 0.028, 120),
]


@pytest.mark.parametrize('spec,p,n', CASES, ids=lambda x: str(x) if not isinstance(x, dict) else
                         f"{x['code']}{x['d']}-{x['noise'].split()[0]}")
@pytest.mark.parametrize('variant', ['default', 'west', 'dynamic'])
def test_pipeline_vs_oracle_on_gpu_samples(spec, p, n, variant):
    """sample (K1) -> decode (K2) -> check (K4) and the fused kernel, against the
    oracle on the errors the GPU drew.  n is BASELINE-size for cfg1/2/3."""
    from localuf_b200 import _engine
    flags = FLAGS[variant]
    code = make_code(spec)
    eng = engine_of(code)
    orc = binding.Oracle.from_code(code)
    seed, shot0 = 99, 12345
    err, syn = eng.sample_host(p, seed, shot0, n)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    corr, growth2b, steps = eng.decode_batch_host(_engine.DEC_UF, flags, syn, want_growth=True)
    ocorr, ogrowth, osteps = orc.decode_batch(0, flags, syn)
    np.testing.assert_array_equal(_engine.unpack_growth(growth2b, code.N_EDGES), ogrowth)
    np.testing.assert_array_equal(steps[:, 0], osteps[:, 0])
    np.testing.assert_array_equal(corr, ocorr)
    # round trip: the correction reproduces the syndrome
    np.testing.assert_array_equal(orc.syndrome(corr), syn)
    logical, fails = eng.check_host(err, corr)
    ological = orc.logical_batch(err, ocorr)
    np.testing.assert_array_equal(logical, ological)
    assert fails == int(ological.sum())
    fused, shots = eng.mc_run_host(_engine.DEC_UF, flags, p, seed, shot0, n)
    assert (fused, shots) == (fails, n)


@pytest.mark.skipif(os.environ.get('LUF_TILE') is not None, reason='already inside the forced-tile run')
@pytest.mark.parametrize('tile', ['cta', 'warp'])
def test_both_tile_shapes_on_every_case(tile):
    """Golden and oracle parity again with the tile shape forced: one thread block
    per shot on the small graphs, one warp per shot on the large ones."""
    import subprocess
    env = dict(os.environ, LUF_TILE=tile)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-x', '-q', '-m', 'gpu',
                        '-k', 'goldens or pipeline', '-p', 'no:cacheprovider'],
                       env=env, cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_sampler_statistics_and_counter_based_stream():
    code = make_code(dict(code='Surface', d=5, noise='circuit-level'))
    eng = engine_of(code)
    g = code.FLAT
    p, n = 0.02, 40000
    err, _ = eng.sample_host(p, 5, 0, n)
    bits = np.unpackbits(err.view(np.uint8), axis=1, bitorder='little')[:, :g.n_edges]
    for c, pc in enumerate(code.NOISE.class_probabilities(p)):
        members = np.flatnonzero(g.edge_class == c)
        k, tot = int(bits[:, members].sum()), n * len(members)
        assert abs(k - tot * pc) < 5 * (tot * pc * (1 - pc)) ** 0.5 + 1, (c, k, tot * pc)
    # per-edge frequencies are uniform inside a class (no position bias from the lane split)
    c0 = np.flatnonzero(g.edge_class == 0)
    freq = bits[:, c0].mean(axis=0)
    pc = code.NOISE.class_probabilities(p)[0]
    assert np.all(np.abs(freq - pc) < 6 * (pc * (1 - pc) / n) ** 0.5)
    # shot k's sample depends only on (seed, k): any split of the range agrees
    a, _ = eng.sample_host(p, 5, 1000, 64)
    np.testing.assert_array_equal(a, err[1000:1064])
    b, _ = eng.sample_host(p, 6, 1000, 64)
    assert not np.array_equal(a, b)
    f_all, _ = eng.mc_run_host(0, 0, p, 5, 0, 3000)
    f_a, _ = eng.mc_run_host(0, 0, p, 5, 0, 1234)
    f_b, _ = eng.mc_run_host(0, 0, p, 5, 1234, 3000 - 1234)
    assert f_all == f_a + f_b


def test_extreme_probabilities_and_empty_batches():
    from localuf_b200 import _engine
    code = make_code(dict(code='Surface', d=5, noise='phenomenological'))
    eng = engine_of(code)
    err, syn = eng.sample_host(0.0, 1, 0, 100)
    assert not err.any() and not syn.any()
    assert eng.mc_run_host(0, 0, 0.0, 1, 0, 5000) == (0, 5000)
    err, syn = eng.sample_host(1.0, 1, 0, 10)
    bits = np.unpackbits(err.view(np.uint8), axis=1, bitorder='little')[:, :code.N_EDGES]
    assert bits.all()
    assert eng.mc_run_host(0, 0, 0.01, 1, 0, 0) == (0, 0)
    corr, _, steps = eng.decode_batch_host(_engine.DEC_UF, 0, np.zeros((3, eng.WD), np.uint32))
    assert not corr.any() and not steps.any()
    # all detectors defective: the maximal syndrome still validates
    full = np.zeros((1, eng.WD), np.uint32)
    D = code.FLAT.n_detectors
    full[0, :D // 32] = 0xFFFFFFFF
    if D % 32:
        full[0, D // 32] = (1 << (D % 32)) - 1
    orc = binding.Oracle.from_code(code)
    corr, _, _ = eng.decode_batch_host(_engine.DEC_UF, 0, full)
    ocorr, _, _ = orc.decode_batch(0, 0, full)
    np.testing.assert_array_equal(corr, ocorr)
    np.testing.assert_array_equal(orc.syndrome(corr), full)


def test_failure_rates_within_binomial_ci():
    """GPU RNG + GPU decode vs exact / oracle-Monte-Carlo failure rates."""
    import localuf_b200 as L
    # repetition d=5: UF corrects weight <= 2, fails on weight >= 3 -> exact rate
    code = L.Repetition(5, 'code capacity')
    eng = engine_of(code)
    n, p = 2_000_000, 0.05
    m, _ = eng.mc_run_host(0, 0, p, 11, 0, n)
    exact = sum(comb(5, k) * p**k * (1 - p)**(5 - k) for k in range(3, 6))
    assert abs(m / n - exact) < 5 * (exact * (1 - exact) / n) ** 0.5
    # surface codes: oracle Monte Carlo with its own (xoshiro) sampler as the reference rate
    for spec, p, n_gpu, n_cpu in [
        (dict(code='Surface', d=9, noise='code capacity'), 0.08, 2_000_000, 400_000),
        (dict(code='Surface', d=5, noise='phenomenological'), 0.03, 2_000_000, 300_000),
        (dict(code='Surface', d=5, noise='circuit-level'), 0.01, 1_000_000, 100_000),
    ]:
        code = make_code(spec)
        eng = engine_of(code)
        orc = binding.Oracle.from_code(code)
        cp = code.NOISE.class_probabilities(p)
        m_gpu, _ = eng.mc_run_host(0, 0, p, 21, 0, n_gpu)
        m_cpu = orc.mc_batch(0, 0, cp, 22, n_cpu, nthreads=len(os.sched_getaffinity(0)))
        r_gpu, r_cpu = m_gpu / n_gpu, m_cpu / n_cpu
        pool = (m_gpu + m_cpu) / (n_gpu + n_cpu)
        sigma = (pool * (1 - pool) * (1 / n_gpu + 1 / n_cpu)) ** 0.5
        assert abs(r_gpu - r_cpu) < 5 * sigma, (spec, r_gpu, r_cpu, sigma)


def test_python_api_drop_in():
    """Decoder.decode / Scheme.run / sim.accuracy.monte_carlo with the reference's shapes."""
    import localuf_b200 as L
    meta, arr = load_golden('uf_sf9_cc_p10')
    code = make_code(meta['spec'])
    g = code.FLAT
    uf = L.decoders.UF(code)
    for s in range(10):
        syndrome = g.bits_to_syndrome(arr['syn_bits'][s])
        uf.reset()
        assert uf.decode(syndrome) is None
        assert uf.correction == g.bits_to_edges(arr['default_corr'][s])
        assert code.get_syndrome(uf.correction) == syndrome
        assert uf.erasure == {g.edges[k] for k in np.flatnonzero(arr['default_growth'][s] == 2)}
        assert {int(v) for v in uf.growth.values()} <= {0, 1, 2}
        error = g.bits_to_edges(arr['err_bits'][s])
        assert code.SCHEME._sim_cycle_given_error(uf, error) == arr['default_logical'][s]
    random.seed(3)
    m1, n1 = code.SCHEME.run(uf, 0.10, 50_000)
    random.seed(3)
    m2, _ = code.SCHEME.run(uf, 0.10, 50_000)
    assert m1 == m2 and n1 == 50_000 and 0 < m1 < 50_000       # random.seed makes runs reproducible
    assert code.SCHEME.run(uf, 0.0, 1000) == (0, 1000)
    phen = L.Surface(5, 'phenomenological')
    assert phen.SCHEME.run(L.decoders.UF(phen), 0.0, 700) == (0, 700.0)
    df = L.sim.accuracy.monte_carlo({3: [(0.05, 2000), (0.1, 2000)], 5: [(0.05, 2000)]},
                                    L.Surface, L.decoders.UF, 'code capacity')
    assert list(df.index) == ['m', 'n'] and list(df.columns.names) == ['d', 'p']
    assert df.shape == (2, 3) and (df.loc['n'] == 2000).all()
    with pytest.raises(ValueError):
        L.decoders.UF(L.Surface(3, 'phenomenological', scheme='frugal'))


def test_graphs_beyond_the_16bit_layout_run_on_the_wide_instance():
    """Beyond the 16-bit layout (N <= 32767, E <= 65533) UF takes the wide instance of the kernels (luf_uf_wide.cuh;
    parity in tests/test_gpu_large.py); Macar / Actis keep their 16-bit cellular-automaton layout and say so."""
    import localuf_b200 as L
    from localuf_b200 import _engine
    big = L.Surface(23, 'circuit-level')          # E > 65533
    eng = _engine.Engine(big, device=0)
    fails, shots = eng.mc_run_host(_engine.DEC_UF, 0, 0.001, 1, 0, 10)
    assert shots == 10 and 0 <= fails <= 10
    with pytest.raises(NotImplementedError, match='16-bit'):
        eng.mc_run_host(_engine.DEC_MACAR, 0, 0.001, 1, 0, 10)
    err, syn = eng.sample_weight_host(3, 1, 0, 10)
    assert np.unpackbits(err.view(np.uint8), axis=1).sum(axis=1).tolist() == [3] * 10
    huge = L.Surface(11, 'phenomenological', scheme='global batch', window_height=11 * 30)     # 36300 detectors
    eng = _engine.Engine(huge, device=0)
    err, syn = eng.sample_host(0.005, 1, 0, 4)
    corr, _, _ = eng.decode_batch_host(_engine.DEC_UF, 0, syn)
    assert corr.shape == err.shape


def test_sampler_groups_of_four_and_unaligned_buffers():
    """k_sample takes four shots per warp with 16-byte stores when the buffers allow: any shot count, any shot
    offset and buffers that are only 4-byte aligned give the rows of one big aligned run."""
    import torch
    from localuf_b200 import _engine
    code = make_code(dict(code='Surface', d=7, noise='phenomenological'))        # W(E) = 29, W(D) = 10: rows of 116 / 40 bytes
    eng = engine_of(code)
    p, seed = 0.03, 17
    err, syn = eng.sample_host(p, seed, 0, 1001)
    for shot0, n in ((0, 1), (3, 2), (5, 3), (6, 5), (997, 4), (10, 7)):
        e, s = eng.sample_host(p, seed, shot0, n)
        np.testing.assert_array_equal(e, err[shot0:shot0 + n])
        np.testing.assert_array_equal(s, syn[shot0:shot0 + n])
    dev = torch.device('cuda', 0)
    n = 37
    ebuf = torch.zeros(n * eng.WE + 1, dtype=torch.int32, device=dev)
    sbuf = torch.zeros(n * eng.WD + 3, dtype=torch.int32, device=dev)
    with torch.cuda.device(dev):
        eng.sample(p, seed, 100, n, ebuf[1:], sbuf[3:])                          # 4-byte aligned only
        torch.cuda.synchronize()
    np.testing.assert_array_equal(ebuf[1:].cpu().numpy().view(np.uint32).reshape(n, eng.WE), err[100:100 + n])
    np.testing.assert_array_equal(sbuf[3:].cpu().numpy().view(np.uint32).reshape(n, eng.WD), syn[100:100 + n])
    assert int(ebuf[0]) == 0 and not sbuf[:3].any()
