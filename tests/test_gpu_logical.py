"""Per-shot parity of the FUSED Monte-Carlo kernels (the headline path), through the C ABI.

``luf_mc_run_logical`` keeps what ``Batch._sim_cycle_given_p`` returns for every shot
(``localuf/_schemes.py:143-155``).  The sampler is counter-based, so ``luf_sample`` with the same
(seed, shot0) exports the very errors the fused kernel drew; the oracle decodes them in canonical
order and its logical bit must equal the fused kernel's on EVERY shot -- for every decoder flag set
that has its own fused instance, on both tile shapes, from low noise up to threshold.  Also here:
the large-distance parity cases (UF d = 21 / 25 staged vs the oracle, Snowflake d = 21 / 25 streams).
"""
import os
import subprocess
import sys

import numpy as np
import pytest

from conftest import ROOT, make_code

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu

CFG3 = dict(code='Surface', d=11, noise='phenomenological')
# (spec, p, n)
UF_CASES = [
    (dict(code='Repetition', d=5, noise='code capacity'), 0.05, 10000),
    (dict(code='Repetition', d=9, noise='phenomenological'), 0.06, 4000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.05, 20000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.10, 20000),
    (CFG3, 0.005, 20000),
    (CFG3, 0.01, 20000),
    (CFG3, 0.02, 8000),
    (CFG3, 0.026, 8000),
    (dict(code='Surface', d=7, noise='circuit-level'), 0.01, 4000),
    (dict(code='Surface', d=13, noise='phenomenological'), 0.026, 1500),
    (dict(code='Surface', d=15, noise='phenomenological'), 0.02, 600),
    (dict(code='Surface', d=17, noise='phenomenological'), 0.026, 200),
]


def _id(x):
    return f"{x['code']}{x['d']}-{x['noise'].split()[0]}" if isinstance(x, dict) else str(x)


def fused_vs_oracle(spec, p, n, flags, decoder=0, seed=20260101, shot0=777):
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    err, syn = eng.sample_host(p, seed, shot0, n)
    ocorr, _, _ = orc.decode_batch(decoder, flags, syn)
    want = orc.logical_batch(err, ocorr)
    logical, fails, shots = eng.mc_logical_host(decoder, flags, p, seed, shot0, n)
    assert shots == n
    bad = np.flatnonzero(logical != want)
    assert bad.size == 0, f'{bad.size} of {n} shots differ, first at {bad[:5]}'
    assert fails == int(want.sum())
    # the count-only entry point launches the same kernels
    assert eng.mc_run_host(decoder, flags, p, seed, shot0, n) == (fails, n)
    return int(want.sum())


@pytest.mark.parametrize('spec,p,n', UF_CASES, ids=_id)
@pytest.mark.parametrize('flags', [0, 2, 4, 12], ids=['default', 'west', 'nodeuf', 'nodebuf'])
def test_fused_uf_logical_bit_per_shot(spec, p, n, flags):
    if flags == 12 and spec['noise'] != 'code capacity':
        pytest.skip('the reference NodeBUF is defined for code-capacity graphs (node_buf.py)')
    fused_vs_oracle(spec, p, n, flags)


@pytest.mark.parametrize('flags', [1, 3], ids=['dynamic', 'dynamic-west'])
def test_fused_dynamic_uf_logical_bit_per_shot(flags):
    fused_vs_oracle(CFG3, 0.02, 4000, flags)
    fused_vs_oracle(dict(code='Surface', d=9, noise='code capacity'), 0.10, 8000, flags)


@pytest.mark.parametrize('spec,p,n', [
    (dict(code='Surface', d=5, noise='phenomenological'), 0.03, 4000),
    (CFG3, 0.01, 4000),                    # block tile (1452 nodes)
    (CFG3, 0.026, 1000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.08, 6000),
], ids=_id)
def test_fused_macar_logical_bit_per_shot(spec, p, n):
    fused_vs_oracle(spec, p, n, 0, decoder=1)


@pytest.mark.skipif(os.environ.get('LUF_TILE') is not None, reason='already inside the forced-tile run')
@pytest.mark.parametrize('tile', ['cta', 'warp'])
def test_fused_logical_on_both_tile_shapes(tile):
    """The same per-shot checks with the tile shape forced (LUF_TILE is read once per process)."""
    env = dict(os.environ, LUF_TILE=tile)
    r = subprocess.run([sys.executable, '-m', 'pytest', os.path.abspath(__file__), '-x', '-q', '-m', 'gpu',
                        '-k', 'fused_uf_logical or dynamic', '-p', 'no:cacheprovider'],
                       env=env, cwd=ROOT, capture_output=True, text=True)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-2000:]


def test_logical_output_is_chunk_independent_and_optional():
    from localuf_b200 import _engine
    code = make_code(CFG3)
    eng = _engine.Engine(code, device=0)
    p, seed = 0.02, 5
    a, fa, _ = eng.mc_logical_host(0, 0, p, seed, 0, 3000)
    b, fb, _ = eng.mc_logical_host(0, 0, p, seed, 1000, 500)
    np.testing.assert_array_equal(a[1000:1500], b)
    assert fa == int(a.sum()) and fb == int(b.sum())
    assert eng.mc_logical_host(0, 0, p, seed, 0, 0)[1:] == (0, 0)


# ------------------------------------------------------------------ large distances
@pytest.mark.parametrize('d,p,n', [(21, 0.026, 300), (25, 0.026, 200), (25, 0.015, 200)])
@pytest.mark.parametrize('variant,flags', [('default', 0), ('west', 2)])
def test_large_distance_uf_vs_oracle(d, p, n, variant, flags):
    """Surface d = 21 / 25 phenomenological near threshold (BASELINE configs[4]): growth, number of
    rounds, corrections and logical bits of the staged pipeline equal the oracle's on the errors the
    GPU drew; the fused kernel's per-shot bit too."""
    from localuf_b200 import _engine
    spec = dict(code='Surface', d=d, noise='phenomenological')
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    seed, shot0 = 4242, 100
    err, syn = eng.sample_host(p, seed, shot0, n)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    corr, growth2b, steps = eng.decode_batch_host(_engine.DEC_UF, flags, syn, want_growth=True)
    ocorr, ogrowth, osteps = orc.decode_batch(0, flags, syn)
    np.testing.assert_array_equal(_engine.unpack_growth(growth2b, code.N_EDGES), ogrowth)
    np.testing.assert_array_equal(steps[:, 0], osteps[:, 0])
    np.testing.assert_array_equal(corr, ocorr)
    np.testing.assert_array_equal(orc.syndrome(corr), syn)
    want = orc.logical_batch(err, ocorr)
    logical, fails = eng.check_host(err, corr)
    np.testing.assert_array_equal(logical, want)
    fused, ffails, _ = eng.mc_logical_host(_engine.DEC_UF, flags, p, seed, shot0, n)
    np.testing.assert_array_equal(fused, want)
    assert ffails == fails == int(want.sum())


@pytest.mark.parametrize('d,p,nstreams', [(21, 0.006, 4), (25, 0.006, 3)])
def test_large_distance_snowflake_streams_vs_oracle(d, p, nstreams):
    """Snowflake on Surface d = 21 / 25 circuit-level frugal windows: per-cycle runtimes, floor
    corrections and logical counts of every stream equal the oracle's."""
    from test_gpu_sf import test_sf_streams_vs_oracle_on_random_errors as run
    run(dict(code='Surface', d=d, noise='circuit-level', kwargs=dict(scheme='frugal')), p, 0, nstreams, 3 * d)


@pytest.mark.parametrize('spec,p', [
    (dict(code='Surface', d=11, noise='phenomenological'), 0.02),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02),
    (dict(code='Surface', d=17, noise='phenomenological'), 0.015),
])
def test_logical_frames_only(spec, p):
    """luf_decode_frames_host: one byte per shot = the parity of the correction on the logical mask
    (_schemes.py:123-129); XOR with the error's parity it is the shot's logical error."""
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    n = 70000 if spec['d'] == 11 else 3000          # more than one chunk of the pipelined host path
    err, syn = eng.sample_host(p, 21, 0, n)
    for dec, flags in ((_engine.DEC_UF, 0), (_engine.DEC_UF, 2), (_engine.DEC_MACAR, 0)):
        frame, steps = eng.decode_frames_host(dec, flags, syn, want_steps=True)
        corr, _, steps2 = eng.decode_batch_host(dec, flags, syn)
        want, _ = eng.check_host(np.zeros_like(corr), corr)
        np.testing.assert_array_equal(frame, want)
        np.testing.assert_array_equal(steps, steps2)
        logical, _ = eng.check_host(err, corr)
        err_par, _ = eng.check_host(err, np.zeros_like(corr))
        np.testing.assert_array_equal(frame ^ err_par, logical)
    ocorr, _, _ = orc.decode_batch(0, 0, syn[:500], want_growth=False)
    frame = eng.decode_frames_host(_engine.DEC_UF, 0, syn[:500])
    np.testing.assert_array_equal(frame ^ orc.logical_batch(err[:500], np.zeros_like(ocorr)), orc.logical_batch(err[:500], ocorr))


@pytest.mark.parametrize('spec,p,n', [
    (dict(code='Surface', d=9, noise='phenomenological'), 0.03, 6000),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.03, 6000),
    (dict(code='Surface', d=17, noise='phenomenological'), 0.027, 1500),
])
def test_hard_shots_every_route_gives_the_same_bits(spec, p, n, monkeypatch):
    """Shots with an odd cluster on both sides (default inclination): one pass with the merge log in global memory +
    k_hard_sort + k_hard_resolve, second pass with the log in shared memory + k_hard_resolve, both with an export
    buffer that holds only a few shots (the rest falls back to the canonical launch), and the canonical launch
    alone -- per-shot bits equal to the oracle's on every route."""
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    err, syn = eng.sample_host(p, 13, 50, n)
    ocorr, _, _ = orc.decode_batch(0, 0, syn, want_growth=False)
    want = orc.logical_batch(err, ocorr)
    for env in ({}, {'LUF_FAST_ONEPASS': '1'}, {'LUF_FAST_ONEPASS': '0'}, {'LUF_FAST_ONEPASS': '1', 'LUF_FAST_XCAP_WORDS': '3000'},
                {'LUF_FAST_ONEPASS': '0', 'LUF_FAST_XCAP_WORDS': '3000'}, {'LUF_FAST_LOG': '0'}):
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        eng2 = _engine.Engine(code, device=0)
        logical, fails, shots = eng2.mc_logical_host(_engine.DEC_UF, 0, p, 13, 50, n)
        np.testing.assert_array_equal(logical, want)
        assert (fails, shots) == (int(want.sum()), n)
        for k in env:
            monkeypatch.delenv(k)
