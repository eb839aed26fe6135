"""C oracle of the forward scheme (oracle/luf_oracle.c: oracle_forward_stream)
vs. goldens exported from the reference (tests/golden/fw_*.npz)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402
from localuf_b200 import _window  # noqa: E402


@pytest.mark.parametrize('name', golden_names('fw_'))
def test_forward_oracle_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    orc = binding.ForwardOracle(code, _window.build_forward(code))
    for key, dec in (('uf', 0), ('macar', 1)):
        steps, commit, m = orc.stream(dec, 0, arr[f'{key}_init'], arr[f'{key}_fresh'])
        np.testing.assert_array_equal(steps, arr[f'{key}_steps'], err_msg=f'{name} {key} steps')
        np.testing.assert_array_equal(commit, arr[f'{key}_commit'], err_msg=f'{name} {key} commit leftover')
        np.testing.assert_array_equal(m, arr[f'{key}_m'], err_msg=f'{name} {key} logical errors')
    assert meta['m']['macar'] == meta['unmodified_macar_m']
