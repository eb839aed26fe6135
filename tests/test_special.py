"""``sim.accuracy.monte_carlo_special`` (localuf/sim/accuracy/monte_carlo.py:162-201): circuit-level errors
decoded on the phenomenological graph.  Goldens (oracle/gen_golden_special.py): the reference's own
``Batch._sim_cycle_given_error`` per shot with canonical-order UF (default / west; west asserted equal to the
unmodified reference while generating) and the unmodified Macar."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402

CASES = [('uf', 0, 0), ('west', 0, 2), ('macar', 1, 0)]


def codes_of(meta):
    import localuf_b200 as L
    return L.Surface(meta['d'], 'circuit-level'), L.Surface(meta['d'], 'phenomenological')


@pytest.mark.parametrize('name', golden_names('sp_'))
def test_oracle_matches_reference(name):
    meta, arr = load_golden(name)
    code, phen = codes_of(meta)
    oc, op = binding.Oracle.from_code(code), binding.Oracle.from_code(phen)
    syn = oc.syndrome(arr['err_bits'])
    for key, dec, flags in CASES:
        corr, _, _ = op.decode_batch(dec, flags, syn)
        np.testing.assert_array_equal(corr, arr[f'{key}_corr'])
        logical = oc.logical_batch(arr['err_bits'], np.zeros_like(arr['err_bits'])) ^ op.logical_batch(np.zeros_like(corr), corr)
        np.testing.assert_array_equal(logical, arr[f'{key}_logical'])


@pytest.mark.gpu
@pytest.mark.parametrize('name', golden_names('sp_'))
def test_device_matches_reference(name):
    import localuf_b200 as L
    from localuf_b200 import _engine
    meta, arr = load_golden(name)
    code, phen = codes_of(meta)
    ec, ep = _engine.Engine(code, device=0), _engine.Engine(phen, device=0)
    oc = binding.Oracle.from_code(code)
    syn = oc.syndrome(arr['err_bits'])
    for key, dec, flags in CASES:
        corr, _, _ = ep.decode_batch_host(dec, flags, syn, want_steps=False)
        np.testing.assert_array_equal(corr, arr[f'{key}_corr'])
        la, _ = ec.check_host(arr['err_bits'], np.zeros_like(arr['err_bits']))
        lb, _ = ep.check_host(np.zeros_like(corr), corr)
        np.testing.assert_array_equal(la ^ lb, arr[f'{key}_logical'])
    # the driver on device-sampled errors: the count equals the oracle's on the same errors
    p, n, seed = meta['p'], 3000, 9
    fails, shots = ep.mc_special(ec, _engine.DEC_UF, 0, p, seed, 0, n, chunk=1024)
    err, syn = ec.sample_host(p, seed, 0, n)
    op = binding.Oracle.from_code(phen)
    corr, _, _ = op.decode_batch(0, 0, syn)
    want = oc.logical_batch(err, np.zeros_like(err)) ^ op.logical_batch(np.zeros_like(corr), corr)
    assert (fails, shots) == (int(want.sum()), n)
    df = L.sim.accuracy.monte_carlo_special([meta['d']], [p, 2 * p], [500, 300], L.Surface, L.decoders.UF)
    assert list(df.loc['n']) == [500.0, 300.0] and df.shape == (2, 2)
    assert 0 <= df.loc['m', (meta['d'], p)] <= 500
    with pytest.raises(ValueError):
        L.sim.accuracy.monte_carlo_special([3], [0.01, 0.02], [10], L.Surface, L.decoders.UF)
