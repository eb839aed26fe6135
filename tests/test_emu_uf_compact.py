"""The compact fast path of the fused UF kernels (localuf_b200/csrc/luf_uf_fast.cuh) on the host emulation:
the very phase source the kernel k_mc_uf_fast runs, lanes of a tile one after the other.

Every shot the path finishes must carry the logical bit of the canonical decode (oracle-checked emulation of
the staged path on the same counter-based noise); a shot it gives up on (HARD: a cluster touches both sides, or
a work list overflowed) is decoded by the canonical kernel on the device, so it may never be wrong, only HARD.
"""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, make_code

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import binding  # noqa: E402
import emu_binding  # noqa: E402

CASES = [
    (dict(code='Repetition', d=5, noise='code capacity'), 0.05, 3000),
    (dict(code='Repetition', d=9, noise='phenomenological'), 0.06, 1500),
    (dict(code='Surface', d=3, noise='code capacity'), 0.12, 3000),
    (dict(code='Surface', d=9, noise='code capacity'), 0.10, 3000),
    (dict(code='Surface', d=5, noise='phenomenological'), 0.03, 3000),
    (dict(code='Surface', d=7, noise='phenomenological', kwargs=dict(window_height=3)), 0.03, 2000),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.01, 2000),
    (dict(code='Surface', d=11, noise='phenomenological'), 0.026, 1000),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.01, 2000),
    (dict(code='Surface', d=7, noise='circuit-level', kwargs=dict(parametrization='standard')), 0.005, 1500),
    (dict(code='Surface', d=13, noise='phenomenological'), 0.02, 300),
]


def canonical_bits(code, p, seed, shot0, n):
    """Logical bit per shot of the canonical decode, and whether a cluster spans both sides."""
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    _, err, syn, corr = emu.mc_uf(0, p, seed, shot0, n)
    ocorr, _, _ = orc.decode_batch(0, 0, syn)
    np.testing.assert_array_equal(corr, ocorr)
    return emu, orc.logical_batch(err, ocorr)


def canonical_west_bits(code, p, seed, shot0, n):
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    _, err, syn, _ = emu.mc_uf(0, p, seed, shot0, n)
    corr, _, _ = orc.decode_batch(0, 2, syn)
    return orc.logical_batch(err, corr)


@pytest.mark.parametrize('spec,p,n', CASES, ids=lambda x: f"{x['code']}{x['d']}-{x['noise'].split()[0]}" if isinstance(x, dict) else str(x))
def test_compact_path_is_never_wrong(spec, p, n):
    code = make_code(spec)
    emu, want = canonical_bits(code, p, 11, 100, n)
    hard_by_tile = {}
    for nl in (8, 16, 32, 64, 256):
        out = emu.mc_uf_compact(p, 11, 100, n, nl=nl, mu=1e6)          # lists that hold everything
        hard = out == 2                                                # HARD = a cluster touches both sides
        np.testing.assert_array_equal(out[~hard], want[~hard])
        hard_by_tile[nl] = hard
    for nl in (8, 16, 64, 256):                                        # which shots are HARD does not depend on the tile
        np.testing.assert_array_equal(hard_by_tile[nl], hard_by_tile[32])
    # the library's list capacities (from the expected number of faults), and starved lists where every list
    # overflows in nearly every shot: the consumers fall back on the bitmaps / growth words -- same bits, same HARD shots
    for mu, nl in ((-1.0, 16), (0.0, 32), (0.0, 8)):
        out = emu.mc_uf_compact(p, 11, 100, n, nl=nl, mu=mu)
        np.testing.assert_array_equal(out == 2, hard_by_tile[32])
        np.testing.assert_array_equal(out[~hard_by_tile[32]], want[~hard_by_tile[32]])
    # west inclination: never HARD, and the bit of the canonical west decode on every shot
    want_west = canonical_west_bits(code, p, 11, 100, n)
    for mu, nl in ((-1.0, 16), (0.0, 32), (1e6, 64)):
        out = emu.mc_uf_compact(p, 11, 100, n, nl=nl, mu=mu, west=True)
        np.testing.assert_array_equal(out, want_west)


def test_compact_path_west_and_default_share_the_fast_result():
    """A shot that the compact path finishes has no cluster on both sides, so the west and the default
    inclination give the same bit (only the canonical kernel of the HARD shots looks at the flag)."""
    code = make_code(dict(code='Surface', d=9, noise='code capacity'))
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    p, n = 0.10, 3000
    _, err, syn, _ = emu.mc_uf(0, p, 5, 0, n)
    out = emu.mc_uf_compact(p, 5, 0, n, nl=32, mu=1e6)
    ok = out != 2
    for flags in (0, 2):
        corr, _, _ = orc.decode_batch(0, flags, syn)
        np.testing.assert_array_equal(out[ok], orc.logical_batch(err, corr)[ok])


def test_compact_path_sampler_is_the_staged_sampler():
    """Tiles of 8 / 16 lanes walk the 32 virtual sampler lanes in turns: same noise as ph_sample, hence the same
    defects -- checked through the number of shots without any defect and the HARD pattern above; here the
    empty and the saturated ends."""
    code = make_code(dict(code='Surface', d=5, noise='phenomenological'))
    emu = emu_binding.Emu(code)
    assert not emu.mc_uf_compact(0.0, 1, 0, 500, nl=8).any()
    out = emu.mc_uf_compact(0.5, 1, 0, 200, nl=16, mu=1e6)
    _, want = canonical_bits(code, 0.5, 1, 0, 200)
    np.testing.assert_array_equal(out[out != 2], want[out != 2])


def test_compact_sampler_equals_the_canonical_sampler():
    code = make_code(dict(code='Surface', d=5, noise='circuit-level'))
    emu = emu_binding.Emu(code)
    _, err, syn, _ = emu.mc_uf(0, 0.02, 5, 40, 300)
    err2, syn2 = emu.sample_compact(0.02, 5, 40, 300)
    np.testing.assert_array_equal(err, err2)
    np.testing.assert_array_equal(syn, syn2)


@pytest.mark.parametrize('d,p,n', [(23, 0.004, 40), (25, 0.006, 30)])
def test_compact_path_on_graphs_beyond_the_16_bit_layout(d, p, n):
    """Surface(23 | 25, 'circuit-level'): 65869 / 85297 edges -- the canonical kernels' 16-bit edge ids do not
    reach; the compact path moves the edge / other-end split of its adjacency entries (17 + 14 bits) and, under
    the west inclination, decodes every shot itself.  Errors from the compact sampler, logical bits against the
    oracle's west decode (equal to the unmodified reference on every shot)."""
    code = make_code(dict(code='Surface', d=d, noise='circuit-level'))
    assert code.N_EDGES > 65533
    emu = emu_binding.Emu(code)
    orc = binding.Oracle.from_code(code)
    err, syn = emu.sample_compact(p, 3, 7, n)
    np.testing.assert_array_equal(orc.syndrome(err), syn)
    corr, _, _ = orc.decode_batch(0, 2, syn)
    want = orc.logical_batch(err, corr)
    for nl, mu in ((256, -1.0), (32, 0.0)):
        out = emu.mc_uf_compact(p, 3, 7, n, nl=nl, mu=mu, west=True)
        np.testing.assert_array_equal(out, want)
    # default inclination: the shots it finishes agree with the canonical default decode
    corr0, _, _ = orc.decode_batch(0, 0, syn)
    out = emu.mc_uf_compact(p, 3, 7, n, nl=128)
    ok = out != 2
    np.testing.assert_array_equal(out[ok], orc.logical_batch(err, corr0)[ok])


@pytest.mark.parametrize('spec,p,n', [c for c in CASES if c[1] >= 0.02 or c[0]['noise'] == 'code capacity'] + [
    (dict(code='Surface', d=7, noise='phenomenological'), 0.04, 3000),
    (dict(code='Surface', d=5, noise='circuit-level'), 0.03, 2000),
    (dict(code='Surface', d=9, noise='phenomenological'), 0.035, 1500),
], ids=lambda x: f"{x['code']}{x['d']}-{x['noise'].split()[0]}" if isinstance(x, dict) else str(x))
def test_hard_clusters_resolved_from_the_merge_log(spec, p, n):
    """Shots with odd clusters on both sides: the merges inside those clusters, replayed in canonical order from
    the merge log (hard_resolve), give the bit of the canonical decode -- on every shot, for every tile width."""
    code = make_code(spec)
    emu, want = canonical_bits(code, p, 11, 100, n)
    hard = emu.mc_uf_compact(p, 11, 100, n, nl=32, mu=1e6) == 2
    for nl, mu in ((32, 1e6), (8, 1e6), (128, -1.0)):
        out, nkeys = emu.mc_uf_compact_log(p, 11, 100, n, nl=nl, mu=mu, lcap=100000, want_keys=True)
        np.testing.assert_array_equal(out, want)
        np.testing.assert_array_equal(nkeys > 0, hard)
    assert hard.sum() > 0 or p < 0.1
    # one-pass mode: the log outside the tile's state, the keys exported as they lie, hard_sort, hard_resolve
    for nl, lcap in ((32, 100000), (128, -1), (64, 256)):
        out, nkeys = emu.mc_uf_compact_log(p, 11, 100, n, nl=nl, lcap=lcap, want_keys=True, onepass=True)
        np.testing.assert_array_equal(out[out != 2], want[out != 2])
        if lcap == 100000:
            np.testing.assert_array_equal(out, want)
            np.testing.assert_array_equal(nkeys > 0, hard)
    # the library's log size: a shot it cannot hold is HARD (2) as before, never wrong
    out = emu.mc_uf_compact_log(p, 11, 100, n, nl=32)
    np.testing.assert_array_equal(out[out != 2], want[out != 2])
    assert (out == 2).mean() <= 0.02 + 0.2 * hard.mean()
    # logs between half full and full sort in place (no room to gather the HARD clusters' entries apart)
    for lcap in (64, 256, 512):
        out = emu.mc_uf_compact_log(p, 11, 100, n, nl=64, lcap=lcap)
        np.testing.assert_array_equal(out[out != 2], want[out != 2])
        entries = emu.last_log_entries
        np.testing.assert_array_equal(out == 2, hard & (entries > lcap))
    # a starved log: every such shot falls back
    out = emu.mc_uf_compact_log(p, 11, 100, n, nl=32, lcap=1)
    np.testing.assert_array_equal(out == 2, hard)
