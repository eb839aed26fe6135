"""NodeUF, NodeBUF and BUF (localuf/decoders/node_uf.py, node_buf.py, buf.py; SURVEY.md section 8
row f4): frontier-node growth and bucket growth through the UF kernels with the LUF_UF_NODE /
LUF_UF_BUCKET flags.  Goldens: oracle/gen_golden_nuf.py (reference
NodeUF in canonical order; growth asserted equal to the unmodified reference)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))

NODE, BUCKET = 4, 8
FLAGS = {'default': NODE, 'dynamic': NODE | 1, 'west': NODE | 2}


def cases():
    """(golden name, flag sets): NodeUF goldens and NodeBUF goldens (bucket flag added)."""
    out = [(n, FLAGS) for n in golden_names('nuf_')]
    out += [(n, {k: v | BUCKET for k, v in FLAGS.items()}) for n in golden_names('nbuf_')]
    out += [(n, {k: (v & ~NODE) | BUCKET for k, v in FLAGS.items()}) for n in golden_names('buf_')]   # BUF
    return out


CASES = cases()
IDS = [c[0] for c in CASES]


@pytest.mark.parametrize('name,FLAGS', CASES, ids=IDS)
def test_oracle_matches_reference(name, FLAGS):
    import binding
    meta, arr = load_golden(name)
    orc = binding.Oracle.from_code(make_code(meta['spec']))
    for variant, flags in FLAGS.items():
        corr, growth, steps = orc.decode_batch(0, flags, arr['syn_bits'])
        np.testing.assert_array_equal(growth, arr[f'{variant}_growth'])
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])


@pytest.mark.parametrize('name,FLAGS', CASES, ids=IDS)
def test_emu_matches_reference(name, FLAGS):
    import emu_binding
    from localuf_b200._engine import unpack_growth
    meta, arr = load_golden(name)
    emu = emu_binding.Emu(make_code(meta['spec']))
    for variant, flags in FLAGS.items():
        for decode in (emu.decode_uf, lambda f, s: emu.decode_uf_cta(f, s, nl=96)):
            corr, growth2b, steps = decode(flags, arr['syn_bits'])
            np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
            np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
            np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])


def test_growth_rule_differs_from_uf_and_equals_macar():
    """SURVEY finding 4: erasure NodeUF == Macar, != UF on some shots (goldens of both families)."""
    import binding
    meta, arr = load_golden('nuf_sf9_cc_p08')
    orc = binding.Oracle.from_code(make_code(meta['spec']))
    _, g_uf, _ = orc.decode_batch(0, 0, arr['syn_bits'])
    _, g_node, _ = orc.decode_batch(0, NODE, arr['syn_bits'])
    _, g_macar, _ = orc.decode_batch(1, 0, arr['syn_bits'])
    assert (g_uf != g_node).any()
    np.testing.assert_array_equal(g_node == 2, g_macar == 2)


@pytest.mark.gpu
@pytest.mark.parametrize('name,FLAGS', CASES, ids=IDS)
def test_gpu_matches_reference(name, FLAGS):
    from localuf_b200 import _engine
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    eng = _engine.Engine(code, device=0)
    for variant, flags in FLAGS.items():
        corr, growth2b, steps = eng.decode_batch_host(_engine.DEC_UF, flags, arr['syn_bits'], want_growth=True)
        np.testing.assert_array_equal(_engine.unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])
        if str(code.SCHEME) == 'batch':
            logical, _ = eng.check_host(arr['err_bits'], corr)
            np.testing.assert_array_equal(logical, arr[f'{variant}_logical'])


@pytest.mark.gpu
def test_gpu_fused_and_class_api():
    import binding
    import localuf_b200 as L
    from localuf_b200 import _engine
    code = L.Surface(9, 'phenomenological')
    dec = L.decoders.NodeUF(code)
    eng = dec.engine
    orc = binding.Oracle.from_code(code)
    p, n = 0.02, 4000
    for flags in (NODE, NODE | 1, NODE | 2, NODE | BUCKET, NODE | BUCKET | 2, BUCKET, BUCKET | 1):
        err, syn = eng.sample_host(p, 3, 50, n)
        ocorr, _, _ = orc.decode_batch(0, flags, syn, want_growth=False)
        want = int(orc.logical_batch(err, ocorr).sum())
        assert eng.mc_run_host(_engine.DEC_UF, flags, p, 3, 50, n) == (want, n)
    m, slender = code.SCHEME.run(dec, p, n, seed=3)
    assert slender == n and m > 0
    m2, _ = code.SCHEME.run(L.decoders.NodeBUF(code), p, n, seed=3)
    m3, _ = code.SCHEME.run(L.decoders.BUF(code), p, n, seed=3)
    assert 0 < m2 < n and 0 < m3 < n
    syndrome = code.get_syndrome(code.make_error(0.03))
    dec.reset(); dec.decode(syndrome)
    assert code.get_syndrome(dec.correction) == syndrome


def test_emu_bucket_mode_with_tiny_work_lists():
    """3-entry lists: more active nodes than list2 holds -> the bucket round looks them up again."""
    import emu_binding
    from localuf_b200._engine import unpack_growth
    emu_binding.use_tiny_lists(True)
    try:
        for name in ('nbuf_sf9_cc_p15', 'nbuf_sf5_cc_p20', 'nuf_sf5_phen_p05', 'buf_sf5_phen_p05', 'buf_sf5_cl_p02'):
            meta, arr = load_golden(name)
            emu = emu_binding.Emu(make_code(meta['spec']))
            flags = BUCKET if name.startswith('buf') else NODE | (BUCKET if name.startswith('nbuf') else 0)
            for decode in (emu.decode_uf, lambda f, s: emu.decode_uf_cta(f, s, nl=64)):
                corr, growth2b, steps = decode(flags, arr['syn_bits'])
                np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr['default_growth'])
                np.testing.assert_array_equal(steps[:, 0], arr['default_rounds'])
                np.testing.assert_array_equal(corr, arr['default_corr'])
    finally:
        emu_binding.use_tiny_lists(False)
