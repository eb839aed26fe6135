"""BASELINE.json's full sizes on the GPU, through size-independent properties:

* the fused kernel and the staged pipeline (sample -> decode -> check on device
  buffers) draw the same counter-based noise, so their failure counts over the
  whole run must be EQUAL, whatever the chunking;
* every correction of a chunk reproduces its syndrome (decode round trip) and
  lies inside the erasure;
* the failure rate agrees with the oracle's own Monte Carlo within 5 sigma.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, make_code

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

pytestmark = pytest.mark.gpu


def staged_failures(eng, decoder, flags, p, seed, n, chunk):
    import torch
    from localuf_b200 import _engine
    dev = torch.device('cuda', eng.device)
    err = torch.empty((chunk, eng.WE), dtype=torch.int32, device=dev)
    syn = torch.empty((chunk, eng.WD), dtype=torch.int32, device=dev)
    corr = torch.empty((chunk, eng.WE), dtype=torch.int32, device=dev)
    fails = torch.zeros(1, dtype=torch.int64, device=dev)
    done = 0
    with torch.cuda.device(dev):
        while done < n:
            k = min(chunk, n - done)
            eng.sample(p, seed, done, k, err, syn)
            eng.decode_batch(decoder, flags, k, syn, corr)
            eng.check(k, err, corr, None, fails)
            done += k
        last = (err[:k].cpu().numpy().view(np.uint32), syn[:k].cpu().numpy().view(np.uint32),
                corr[:k].cpu().numpy().view(np.uint32))
    return int(fails.item()), last


@pytest.mark.parametrize('spec,p,n,chunk', [
    (dict(code='Repetition', d=5, noise='code capacity'), 0.05, 10_000, 4096),           # configs[0]
    (dict(code='Surface', d=9, noise='code capacity'), 0.05, 10_000_000, 1 << 20),      # configs[1]: 10^7 shots
    (dict(code='Surface', d=11, noise='phenomenological'), 0.01, 4_000_000, 1 << 19),   # configs[2]
], ids=['cfg1', 'cfg2', 'cfg3'])
def test_fused_equals_staged_at_full_size(spec, p, n, chunk):
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    seed = 20260101
    fused, shots = eng.mc_run_host(_engine.DEC_UF, 0, p, seed, 0, n)
    assert shots == n
    staged, (err, syn, corr) = staged_failures(eng, _engine.DEC_UF, 0, p, seed, n, chunk)
    assert fused == staged
    # round trip on the last chunk: correction -> syndrome; logical check re-derived by the oracle
    np.testing.assert_array_equal(orc.syndrome(corr), syn)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    # any split of the shot range gives the same total
    a, _ = eng.mc_run_host(_engine.DEC_UF, 0, p, seed, 0, n // 3)
    b, _ = eng.mc_run_host(_engine.DEC_UF, 0, p, seed, n // 3, n - n // 3)
    assert a + b == fused
    # rate vs the oracle's own Monte Carlo
    n_cpu = min(n, 400_000)
    m_cpu = orc.mc_batch(0, 0, code.NOISE.class_probabilities(p), 5, n_cpu, nthreads=len(os.sched_getaffinity(0)))
    pool = (fused + m_cpu) / (n + n_cpu)
    sigma = (max(pool * (1 - pool), 1e-9) * (1 / n + 1 / n_cpu)) ** 0.5
    assert abs(fused / n - m_cpu / n_cpu) < 5 * sigma + 1e-7


def test_macar_fused_equals_staged_and_steps_sum():
    """configs[2] with Macar: failures of the fused kernel == staged pipeline; (tSV, tBP) sums too."""
    import torch
    from localuf_b200 import _engine
    code = make_code(dict(code='Surface', d=11, noise='phenomenological'))
    eng = _engine.Engine(code, device=0)
    p, seed, n = 0.01, 77, 200_000
    fused, shots = eng.mc_run_host(_engine.DEC_MACAR, 0, p, seed, 0, n)
    sums = eng.last_step_sums
    err, syn = eng.sample_host(p, seed, 0, n)
    corr, _, steps = eng.decode_batch_host(_engine.DEC_MACAR, 0, syn)
    _, fails = eng.check_host(err, corr)
    assert (fused, shots) == (fails, n)
    assert sums == (int(steps[:, 0].sum()), int(steps[:, 1].sum()))
    orc = binding.Oracle.from_code(code)
    np.testing.assert_array_equal(orc.syndrome(corr), syn)


@pytest.mark.parametrize('spec,p,flags,n,chunk', [
    (dict(code='Surface', d=11, noise='phenomenological'), 0.026, 0, 2_000_000, 1 << 19),   # 13 % of the shots take the canonical redo
    (dict(code='Surface', d=11, noise='phenomenological'), 0.02, 2, 1_000_000, 1 << 19),    # west inclination
    (dict(code='Surface', d=9, noise='code capacity'), 0.10, 2, 2_000_000, 1 << 20),
    (dict(code='Surface', d=7, noise='circuit-level'), 0.01, 0, 1_000_000, 1 << 19),
    (dict(code='Surface', d=17, noise='phenomenological'), 0.02, 0, 200_000, 1 << 16),      # one thread block per shot
    (dict(code='Surface', d=17, noise='phenomenological'), 0.03, 2, 60_000, 1 << 15),
    (dict(code='Repetition', d=9, noise='phenomenological'), 0.06, 0, 2_000_000, 1 << 20),
    (dict(code='Surface', d=9, noise='phenomenological'), 0.02, 4, 1_000_000, 1 << 19),     # NodeUF
    (dict(code='Surface', d=9, noise='code capacity'), 0.10, 12, 1_000_000, 1 << 19),       # NodeBUF
], ids=['cfg3-threshold', 'cfg3-west', 'sf9cc-west', 'sf7cl', 'sf17-block', 'sf17-block-west', 'rep9', 'nodeuf', 'nodebuf'])
def test_fused_order_free_path_equals_staged(spec, p, flags, n, chunk):
    """The fused kernels of static UF merge a round's edges in any order and replace the peel by a pass over
    the defects; shots in which a cluster spans both sides are decoded again in canonical
    order.  Same seed, same shots: the failure count equals the staged pipeline's (canonical merge, full
    peel, luf_check) -- near threshold, with the west inclination, on both tile shapes."""
    from localuf_b200 import _engine
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    seed = 77
    fused, shots = eng.mc_run_host(_engine.DEC_UF, flags, p, seed, 0, n)
    staged, _ = staged_failures(eng, _engine.DEC_UF, flags, p, seed, n, chunk)
    assert shots == n and fused == staged and fused > 0


@pytest.mark.parametrize('spec,p,n,chunk', [
    (dict(code='Surface', d=7, noise='code capacity'), 0.10, 500_000, 1 << 18),
    (dict(code='Surface', d=17, noise='phenomenological'), 0.02, 40_000, 1 << 15),          # one thread block per shot
], ids=['warp', 'block'])
def test_fused_kernel_on_a_graph_that_fails_the_parity_check(spec, p, n, chunk):
    """luf_graph_create takes any luf_graph_desc.  With a logical mask that is not "first endpoint is a west
    boundary node" (one bulk edge more) the parity-only peel of the fused kernels would count the wrong bit;
    pack_graph refuses it for such a graph (parity_ok, csrc/luf_host.h) and every shot takes the full peel,
    so the fused count still equals the staged pipeline's (whose luf_check reads the same mask) -- and differs
    from the count under the code's own mask."""
    import dataclasses
    import types
    from localuf_b200 import _engine
    code = make_code(spec)
    flat = code.FLAT
    mask = np.array(flat.west_edge_mask, dtype=np.uint32, copy=True)
    e = next(i for i in range(flat.n_edges // 2, flat.n_edges) if not (int(mask[i >> 5]) >> (i & 31)) & 1)
    mask[e >> 5] |= np.uint32(1 << (e & 31))
    odd = types.SimpleNamespace(FLAT=dataclasses.replace(flat, west_edge_mask=mask), NOISE=code.NOISE)
    eng, eng0 = _engine.Engine(odd, device=0), _engine.Engine(code, device=0)
    seed = 31
    fused, shots = eng.mc_run_host(_engine.DEC_UF, 0, p, seed, 0, n)
    staged, _ = staged_failures(eng, _engine.DEC_UF, 0, p, seed, n, chunk)
    plain, _ = eng0.mc_run_host(_engine.DEC_UF, 0, p, seed, 0, n)
    assert shots == n and fused == staged and fused > 0
    assert fused != plain
