"""Kernel phase logic (localuf_b200/csrc/luf_core.cuh compiled for the host,
lanes run sequentially) vs. reference goldens and the C oracle.  CPU-only: the
real kernels are checked on the GPU by tests/test_gpu_*.py."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import binding  # noqa: E402
import emu_binding  # noqa: E402
from localuf_b200._engine import unpack_growth  # noqa: E402

FLAGS = {'default': 0, 'dynamic': 1, 'west': 2}


@pytest.mark.parametrize('name', golden_names('uf_'))
def test_emu_decode_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    for variant, flags in FLAGS.items():
        corr, growth2b, steps = emu.decode_uf(flags, arr['syn_bits'])
        np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])


@pytest.mark.parametrize('name', golden_names('uf_'))
@pytest.mark.parametrize('nl', [64, 448])
def test_emu_block_tile_matches_reference(name, nl):
    """The block-per-shot instance of the phases (luf::uf_cta; tiles of 64 and 448 lanes)."""
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    for variant, flags in FLAGS.items():
        corr, growth2b, steps = emu.decode_uf_cta(flags, arr['syn_bits'], nl=nl)
        np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
        np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
        np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])


@pytest.mark.parametrize('spec,p', [
    (dict(code='Repetition', d=5, noise='code capacity'), 0.05),
    (dict(code='Surface', d=9, noise='code capacity'), 0.10),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.05),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.02),
])
def test_emu_fused_mc_vs_oracle(spec, p):
    """Fused sample -> decode -> check: the oracle re-derives syndrome,
    correction and logical outcome from the sampled errors."""
    code = make_code(spec)
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    n = 300
    fails, err, syn, corr = emu.mc_uf(0, p, seed=1234, shot0=17, nshots=n)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    ocorr, _, _ = orc.decode_batch(0, 0, syn, want_growth=False)
    np.testing.assert_array_equal(ocorr, corr)
    assert fails == int(orc.logical_batch(err, ocorr).sum())
    # sampler: flip frequency of every probability class within 5 sigma
    g = code.FLAT
    bits = np.unpackbits(err.view(np.uint8), axis=1, bitorder='little')[:, :g.n_edges]
    for c, pc in enumerate(code.NOISE.class_probabilities(p)):
        members = np.flatnonzero(g.edge_class == c)
        k, tot = int(bits[:, members].sum()), n * len(members)
        assert abs(k - tot * pc) < 5 * (tot * pc * (1 - pc)) ** 0.5 + 1
    assert not bits[:, g.edge_class == 255].any()


def test_emu_sampler_extremes():
    code = make_code(dict(code='Surface', d=3, noise='phenomenological'))
    emu = emu_binding.Emu(code)
    _, err, _, _ = emu.mc_uf(0, 0.0, 1, 0, 20)
    assert not err.any()
    _, err, syn, _ = emu.mc_uf(0, 1.0, 1, 0, 20)
    bits = np.unpackbits(err.view(np.uint8), axis=1, bitorder='little')[:, :code.N_EDGES]
    assert bits.all()


@pytest.mark.parametrize('spec', [
    dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='forward')),
    dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='forward')),
    dict(code='Surface', d=4, noise='circuit-level', kwargs=dict(scheme='forward', commit_height=2, buffer_height=3)),
    dict(code='Surface', d=5, noise='circuit-level'),
    dict(code='Surface', d=7, noise='code capacity'),
], ids=lambda s: f"{s['noise'].split()[0]}-{s.get('kwargs', {}).get('scheme', 'batch')}")
def test_emu_decode_random_syndromes_vs_oracle(spec):
    """Dense random errors on every edge of the graph (also the non-fresh ones of a
    forward window, whose future-boundary nodes have degree > 1): kernel phases == oracle."""
    code = make_code(spec)
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    g = code.FLAT
    rng = np.random.default_rng(11)
    n = 1500
    pad = np.zeros((n, 32 * emu.WE), np.uint8)
    pad[:, :g.n_edges] = rng.random((n, g.n_edges)) < rng.choice([0.01, 0.03, 0.08], size=(n, 1))
    syn = orc.syndrome(np.packbits(pad, axis=1, bitorder='little').view(np.uint32))
    for flags in (0, 1, 2):
        corr, growth2b, steps = emu.decode_uf(flags, syn)
        ocorr, ogrowth, osteps = orc.decode_batch(0, flags, syn)
        np.testing.assert_array_equal(unpack_growth(growth2b, g.n_edges), ogrowth)
        np.testing.assert_array_equal(steps[:, 0], osteps[:, 0])
        np.testing.assert_array_equal(corr, ocorr)


def test_emu_overflow_paths_with_tiny_work_lists():
    """Same goldens through a build whose work lists hold 3 entries: the in-place
    growth of overflowing active nodes, the serial ascending-order merge and the
    bitmap fallback of the tree roots all run."""
    emu_binding.use_tiny_lists(True)
    try:
        for name in ('uf_sf9_cc_p10', 'uf_sf7_cc_p25', 'uf_sf5_phen_p08', 'uf_sf5_cl_p03', 'uf_sf4_cl_fwd_p03'):
            meta, arr = load_golden(name)
            emu = emu_binding.Emu(make_code(meta['spec']))
            for variant, flags in FLAGS.items():
                corr, growth2b, steps = emu.decode_uf(flags, arr['syn_bits'])
                np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
                np.testing.assert_array_equal(steps[:, 0], arr[f'{variant}_rounds'])
                np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])
                corr, growth2b, steps = emu.decode_uf_cta(flags, arr['syn_bits'], nl=96)
                np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
                np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])
                corr, growth2b, steps = emu.decode_uf_wide(flags, arr['syn_bits'], nl=96)      # luf::uf_wide, lists of 5 / 3
                np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{variant}_growth'])
                np.testing.assert_array_equal(corr, arr[f'{variant}_corr'])
    finally:
        emu_binding.use_tiny_lists(False)


@pytest.mark.parametrize('spec,p,flags,n', [
    (dict(code='Surface', d=5, noise='code capacity'), 0.12, 0, 4000),
    (dict(code='Surface', d=5, noise='code capacity'), 0.12, 2, 4000),       # west inclination
    (dict(code='Surface', d=5, noise='code capacity'), 0.12, 1, 4000),       # dynamic
    (dict(code='Surface', d=9, noise='code capacity'), 0.10, 0, 2000),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.04, 0, 2000),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.04, 4, 1500),    # NodeUF
    (dict(code='Surface', d=5, noise='phenomenological'), 0.04, 8, 1500),    # BUF
    (dict(code='Surface', d=7, noise='phenomenological'), 0.025, 0, 1500),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02, 0, 1500),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02, 3, 1500),
    (dict(code='Repetition', d=7, noise='phenomenological'), 0.08, 0, 3000),
    (dict(code='Repetition', d=5, noise='code capacity'), 0.2, 0, 4000),
])
@pytest.mark.parametrize('nl', [32, 128])
def test_emu_parity_only_peel_equals_full_peel(spec, p, flags, n, nl):
    """The fused kernels' shortcut (uf_peel_parity: every defect adds 1 iff its cluster's boundary node is a
    west one; shots with a cluster that spans both sides take the full peel) gives the logical outcome of
    the full peel on every shot."""
    code = make_code(spec)
    emu = emu_binding.Emu(code)
    full, fast, differ, mixed = emu.mc_uf_parity(flags, p, 91, n, nl=nl)
    assert differ == 0 and full == fast
    assert full > 0
    if flags == 0 and spec['code'] == 'Surface' and spec['d'] == 5 and spec['noise'] == 'code capacity':
        assert 0 < mixed < n                     # both branches are exercised


def test_emu_parity_peel_guard():
    """pack_graph admits the parity-only peel for the graphs of Surface / Repetition and refuses it for a
    graph whose logical mask is not "first endpoint is a west boundary node"; a refused graph takes the
    full peel on every shot, so the fused path still equals the staged one."""
    for spec in (dict(code='Surface', d=5, noise='code capacity'), dict(code='Surface', d=5, noise='phenomenological'),
                 dict(code='Surface', d=3, noise='circuit-level'), dict(code='Repetition', d=5, noise='phenomenological')):
        assert emu_binding.Emu(make_code(spec)).parity_ok() == 1
    emu = emu_binding.Emu(make_code(dict(code='Surface', d=5, noise='code capacity')))
    mask = emu._keep['west_edge_mask']
    e = next(i for i in range(emu.flat.n_edges) if not (mask[i >> 5] >> (i & 31)) & 1)
    mask[e >> 5] |= np.uint32(1 << (e & 31))         # another logical operator: one bulk edge more
    assert emu.parity_ok() == 0
    full, fast, differ, mixed = emu.mc_uf_parity(0, 0.12, 5, 3000)
    assert differ == 0 and full == fast and full > 0


@pytest.mark.parametrize('spec,p,flags,n', [
    (dict(code='Surface', d=5, noise='code capacity'), 0.12, 0, 4000),
    (dict(code='Surface', d=5, noise='code capacity'), 0.12, 2, 4000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.10, 0, 2000),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.04, 0, 2000),
    (dict(code='Surface', d=7, noise='phenomenological'), 0.025, 2, 1000),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.02, 0, 200),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02, 0, 1500),
    (dict(code='Repetition', d=7, noise='phenomenological'), 0.08, 0, 3000),
    (dict(code='Repetition', d=5, noise='code capacity'), 0.2, 0, 4000),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.04, 4, 1500),     # NodeUF
    (dict(code='Surface', d=7, noise='code capacity'), 0.10, 6, 2000),        # NodeUF, west
    (dict(code='Surface', d=7, noise='code capacity'), 0.10, 12, 2000),       # NodeBUF
    (dict(code='Surface', d=5, noise='circuit-level'), 0.02, 14, 1000),       # NodeBUF, west
])
@pytest.mark.parametrize('nl', [32, 128])
def test_emu_order_free_merge_equals_canonical(spec, p, flags, n, nl):
    """uf_validate_fast (merges of a round in any order) + ph_defect_sides against the canonical decode with
    the full peel: same logical outcome and the same number of growth rounds on every shot; shots with a
    cluster that holds boundary nodes of both sides are flagged (the kernel decodes those canonically)."""
    code = make_code(spec)
    emu = emu_binding.Emu(code)
    canon, fast, differ, mixed, rounds_differ = emu.mc_uf_fast(flags, p, 17, n, nl=nl)
    assert differ == 0 and canon == fast and rounds_differ == 0
    assert canon > 0


def test_emu_order_free_merge_with_tiny_lists():
    emu_binding.use_tiny_lists(True)
    try:
        emu = emu_binding.Emu(make_code(dict(code='Surface', d=5, noise='phenomenological')))
        for nl in (32, 64):
            canon, fast, differ, mixed, rounds_differ = emu.mc_uf_fast(0, 0.04, 3, 1500, nl=nl)
            assert differ == 0 and rounds_differ == 0 and canon > 0
    finally:
        emu_binding.use_tiny_lists(False)


@pytest.mark.parametrize('spec,p,n', [
    (dict(code='Surface', d=9, noise='code capacity'), 0.12, 3000),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.03, 600),
    (dict(code='Surface', d=7, noise='circuit-level'), 0.012, 1000),
    (dict(code='Repetition', d=9, noise='phenomenological'), 0.08, 2000),
], ids=['sf9cc', 'sf11phen', 'sf7cl', 'rep9'])
def test_no_peel_needed_when_boundary_nodes_are_leaves(spec, p, n):
    """Every boundary node of a Surface / Repetition graph has one edge.  The peel roots a cluster's tree at
    cluster.boundary (uf.py:314-324); the cluster's other boundary nodes are leaves and no defects, so their
    edges never enter the correction (uf.py:305-312), and the root's edge does iff the cluster is odd.  After
    the merges in canonical order the pass over the defects (uf_peel_parity with g.leaf_bnd) therefore gives
    the full peel's west parity on EVERY shot -- also on those with a cluster on both sides, for every flag
    set and both tile shapes; the fused kernels never build the forest on such graphs."""
    emu = emu_binding.Emu(make_code(spec))
    for flags in (0, 1, 2, 3, 4, 6) + ((8, 12) if spec['noise'] == 'code capacity' else ()):
        for nl in (32, 128):
            full, sides, differ, mixed = emu.mc_uf_sides(flags, p, 7, n, nl=nl)
            assert differ == 0 and full == sides
            assert mixed > 0 or (flags & 1)          # (dynamic UF never joins two boundary clusters)
