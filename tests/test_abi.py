"""The C-ABI library loads without a GPU and exports every symbol that
include/luf_b200.h declares; without a device every entry point fails loudly."""
import ctypes
import os
import re

import pytest

from conftest import ROOT


def declared_symbols():
    text = open(os.path.join(ROOT, 'include', 'luf_b200.h')).read()
    text = re.sub(r'/\*.*?\*/', '', text, flags=re.S)
    return sorted(set(re.findall(r'\b(luf_[a-z_0-9]+)\s*\(', text)))


@pytest.fixture(scope='module')
def lib():
    from localuf_b200 import build
    return ctypes.CDLL(build.build())


def test_every_declared_symbol_is_exported(lib):
    names = declared_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(lib, n), f'{n} declared in include/luf_b200.h but not exported'


def test_no_cpu_fallback(lib):
    import localuf_b200 as L
    from localuf_b200 import _engine
    if lib.luf_device_count() > 0:
        pytest.skip('a GPU is present')
    code = L.Surface(3, 'code capacity')
    with pytest.raises(_engine.EngineUnavailable):
        _engine.Engine(code)
    with pytest.raises(_engine.EngineUnavailable):
        code.SCHEME.run(L.decoders.UF(code), 0.1, 10)
    # the product package never touches the oracle
    import subprocess, sys
    out = subprocess.run(['grep', '-rIl', '-e', 'oracle', os.path.join(ROOT, 'localuf_b200')],
                         capture_output=True, text=True).stdout.split()
    offenders = [f for f in out if f.endswith(('.py', '.cu', '.cuh', '.h')) and
                 any(('import' in l or 'include' in l or 'CDLL' in l) and 'oracle' in l
                     for l in open(f, errors='ignore'))]
    assert offenders == []


def test_no_cpu_fallback_for_scores_and_streams(lib):
    """The confidence scores and the streaming decoder have no host path either."""
    import localuf_b200 as L
    from localuf_b200 import _dcs, _engine
    if lib.luf_device_count() > 0:
        pytest.skip('a GPU is present')
    code = L.Surface(3, 'code capacity')
    uf = L.decoders.UF(code)
    with pytest.raises(_engine.EngineUnavailable):
        uf.swim_distance()
    with pytest.raises(_engine.EngineUnavailable):
        _dcs.DcsEngine(_dcs.flat_tables(code))
    frugal = L.Surface(3, 'phenomenological', scheme='frugal')
    sf = L.decoders.Snowflake(frugal)
    with pytest.raises(_engine.EngineUnavailable):
        sf.decode(set(), metrics=('swim_distance',))
    with pytest.raises(_engine.EngineUnavailable):
        frugal.SCHEME.run(sf, 0.01, 1, metrics=('min_defect_height',))


def test_bad_arguments_fail_before_any_device_work(lib):
    """Argument checks of the newer entry points run first: NULL handles and descriptions are rejected with
    LUF_ERR_INVALID (-1) and a message, with or without a GPU."""
    lib.luf_last_error.restype = ctypes.c_char_p
    out = ctypes.c_void_p()
    assert lib.luf_dcs_create(None, 0, ctypes.byref(out)) == -1
    assert b'luf_dcs_create' in lib.luf_last_error()
    assert lib.luf_dcs_run(None, ctypes.c_int64(1), None, ctypes.c_int64(1), 1, None, None, None, None) == -1
    assert lib.luf_dcs_run_host(None, ctypes.c_int64(1), None, ctypes.c_int64(1), 1, None, None, None) == -1
    assert lib.luf_stream_decode_metrics(None, 0, ctypes.c_int64(1), 1, None, None, None, None, None, None, None, None) == -1
    assert lib.luf_stream_mc_metrics(None, 0, None, ctypes.c_uint64(0), ctypes.c_uint64(0), ctypes.c_int64(1), 1,
                                     None, None, None, None, None) == -1
    assert lib.luf_stream_growth_words(None) == 0
    assert lib.luf_dcs_destroy(None) == 0
