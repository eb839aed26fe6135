"""C oracle of Snowflake + frugal scheme (oracle/luf_oracle_snowflake.c) vs.
golden vectors exported from the unmodified reference (tests/golden/sf_*.npz)."""
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'oracle'))
import binding  # noqa: E402
from localuf_b200 import _window  # noqa: E402


@pytest.mark.parametrize('name', golden_names('sf_'))
def test_sf_oracle_matches_reference(name):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    t = _window.build(code)
    kw = meta['decoder_kwargs']
    orc = binding.StreamOracle(t, merger=kw.get('merger', 'fast'), schedule=kw.get('schedule', '2:1'))
    fresh = binding.code_bits_to_local(t, arr['fresh_bits'], 0)
    rt_all, rt_unr, floor, m = orc.stream(fresh)
    np.testing.assert_array_equal(rt_all, arr['runtime_all'], err_msg=f'{name} runtime (all)')
    np.testing.assert_array_equal(rt_unr, arr['runtime_unrooting'], err_msg=f'{name} runtime (unrooting)')
    np.testing.assert_array_equal(floor, binding.code_bits_to_local(t, arr['floor_bits'], t.h - 1),
                                  err_msg=f'{name} floor corrections')
    np.testing.assert_array_equal(m, arr['m_cycle'], err_msg=f'{name} logical errors per cycle')
    assert int(m.sum()) == meta['m'] == meta['unmodified_m']
