"""Forward-scheme kernel phases (localuf_b200/csrc/luf_fw.cuh compiled for the
host) vs. goldens exported from the reference.  CPU-only."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import emu_binding  # noqa: E402


@pytest.mark.parametrize('name', golden_names('fw_'))
def test_emu_forward_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    for key, dec in (('uf', 0), ('macar', 1)):
        steps, commit, m = emu.forward_stream(code, dec, 0, arr[f'{key}_init'], arr[f'{key}_fresh'])
        np.testing.assert_array_equal(steps, arr[f'{key}_steps'], err_msg=f'{name} {key} steps')
        np.testing.assert_array_equal(commit, arr[f'{key}_commit'], err_msg=f'{name} {key} commit leftover')
        np.testing.assert_array_equal(m, arr[f'{key}_m'], err_msg=f'{name} {key} logical errors')


@pytest.mark.parametrize('name', golden_names('fw_'))
@pytest.mark.parametrize('nl', [32, 160])
def test_emu_forward_wide_matches_reference(name, nl):
    """The forward scheme on the wide instance (luf_fw_wide.cuh: 32-bit tables, block tile), UF."""
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    emu = emu_binding.Emu(code)
    steps, commit, m = emu.forward_stream_wide(code, 0, arr['uf_init'], arr['uf_fresh'], nl=nl)
    np.testing.assert_array_equal(steps, arr['uf_steps'])
    np.testing.assert_array_equal(commit, arr['uf_commit'])
    np.testing.assert_array_equal(m, arr['uf_m'])
