"""Macar / Actis kernel phases (localuf_b200/csrc/luf_ca.cuh compiled for the
host) vs. goldens exported from the unmodified reference.  CPU-only."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import emu_binding  # noqa: E402
from localuf_b200._engine import unpack_growth  # noqa: E402

DECODERS = {'macar': (1, 0), 'actis': (2, 0), 'actis_unopt': (2, 4)}


@pytest.mark.parametrize('name', golden_names('ca_'))
def test_emu_ca_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    for key in meta['decoders']:
        dec, flags = DECODERS[key]
        corr, growth2b, steps = emu.decode_ca(dec, flags, arr['syn_bits'])
        np.testing.assert_array_equal(steps, arr[f'{key}_steps'], err_msg=f'{name} {key} (tSV, tBP)')
        np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{key}_growth'],
                                      err_msg=f'{name} {key} erasure after validate')
        np.testing.assert_array_equal(corr, arr[f'{key}_corr'], err_msg=f'{name} {key} correction')


def test_emu_macar_bitmap_path_with_tiny_node_list():
    """A build whose touched-node list holds 4 entries: Macar sweeps the bitmap instead."""
    emu_binding.use_tiny_lists(True)
    try:
        for name in ('ca_sf5_cc_p08', 'ca_sf5_phen_p03', 'ca_sf5_cl_p01'):
            meta, arr = load_golden(name)
            emu = emu_binding.Emu(make_code(meta['spec']))
            corr, growth2b, steps = emu.decode_ca(1, 0, arr['syn_bits'])
            np.testing.assert_array_equal(steps, arr['macar_steps'])
            np.testing.assert_array_equal(corr, arr['macar_corr'])
    finally:
        emu_binding.use_tiny_lists(False)


@pytest.mark.parametrize('nl', [64, 256])
@pytest.mark.parametrize('name', golden_names('ca_'))
def test_emu_ca_block_tile_matches_reference(name, nl):
    """The block-per-shot instance of the decode driver (luf::ca_cta; tiles of 64 and 256 lanes)."""
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    for key in meta['decoders']:
        dec, flags = DECODERS[key]
        corr, growth2b, steps = emu.decode_ca(dec, flags, arr['syn_bits'], nl=nl)
        np.testing.assert_array_equal(steps, arr[f'{key}_steps'], err_msg=f'{name} {key} (tSV, tBP)')
        np.testing.assert_array_equal(unpack_growth(growth2b, meta['E']), arr[f'{key}_growth'])
        np.testing.assert_array_equal(corr, arr[f'{key}_corr'], err_msg=f'{name} {key} correction')
