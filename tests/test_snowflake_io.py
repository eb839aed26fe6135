"""Snowflake's bit-string I/O (``Snowflake.generate_output`` / ``Snowflake.decode``,
reference ``snowflake/main.py:216-281,656-883``): syndrome vectors in, floor
vectors out.  Checked against goldens exported from the unmodified reference
(``oracle/gen_golden_io.py``) and against the input/output pairs printed in the
reference's ``demo_notebooks/snowflake_customization.ipynb`` (cells 5 and 7).
CPU: the kernel phases compiled for the host stand in for the device call;
GPU: the product path through the C ABI."""
import csv
import os
import sys

import numpy as np
import pytest

from conftest import ROOT, load_golden, make_code, golden_names

sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import localuf_b200 as L  # noqa: E402
from localuf_b200 import _engine  # noqa: E402
from localuf_b200.decoders._bitstrings import Bitstrings  # noqa: E402

NOTEBOOK_REP5_IN = ['1001', '0101', '0100', '0010', '1000', '0110', '0100', '1000', '0010', '0000', '0000', '1000',
                    '0110', '0100', '1000', '0010', '0000', '1001', '0101', '0100', '0010', '1000', '0110', '0100',
                    '1000', '0010', '0000', '0000', '1000', '0010']
NOTEBOOK_REP5_OUT = ['000000000', '000000000', '000000000', '000000000', '000000000', '000110000', '010000000',
                     '000000000', '001000000', '001010000', '010000000', '000000000', '000010000', '001100001',
                     '000000010', '000000000', '100000000', '010001100', '000000000', '000010000', '001000000',
                     '001000000', '000110011', '010000000', '000000000', '001000000', '001010000', '010000000',
                     '000000000', '000010000', '001100001', '000000010', '000000000', '000010000', '001100001']
NOTEBOOK_SF3_IN = ['100100', '010100', '010001', '001000']
NOTEBOOK_SF3_OUT = ['0000000000000000000000000000', '0000000000000000000000000000', '0000000000000000000000000000',
                    '0000000000000000000000001100', '0000010100000000000000000000', '0000000000000000000000000000',
                    '0000000000000000000000100000']


def on_emulation(dec):
    """Route the decoder's device call to the host build of the same kernel phases."""
    import emu_binding
    emu = emu_binding.EmuStream(dec.CODE, dec._tables)          # the decoder's own neighbour order

    def run(rows, drops=None):
        if drops is not None and any(drops):
            rt, floor, _, _, _ = emu.stream_metrics(dec._flags | _engine.SF_SYNDROME, np.stack(rows),
                                                    drop_only=np.array(drops, np.uint8))
            return rt, floor
        rt, floor, _ = emu.stream(dec._flags | _engine.SF_SYNDROME, np.stack(rows))
        return rt, floor
    dec._run_cycles = run
    return dec


def strings(bits):
    return [''.join(map(str, row)) for row in bits]


def check_case(name, prepare):
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    dec = prepare(L.decoders.Snowflake(code, **meta['decoder_kwargs']))
    index = {e: k for k, e in enumerate(code.EDGES)}
    assert [index[e] for e in dec._LOWEST_EDGES] == arr['lowest_edges'].tolist()
    vectors, floors = strings(arr['vectors']), strings(arr['floors'])
    assert dec.generate_output(list(vectors)) == floors
    assert dec.floor_history == floors
    dec.reset()
    sets = dec._BITSTRING_CONVERTER.syndrome_vectors_to_sets(vectors)
    assert dec.generate_output(sets) == floors
    # cycle by cycle; runtimes in all three modes and the floor vector logged before each drop
    for mode in ('all', 'merging', 'unrooting'):
        dec.reset()
        dec.init_history()
        n = min(len(sets), 12)
        got = [dec.decode(s, time_only=mode, log_floor_history=True) for s in sets[:n]]
        assert got == arr[f'runtime_{mode}'][:n].tolist(), mode
        assert dec.floor_history == floors[:n]


@pytest.mark.parametrize('name', golden_names('io_'))
def test_emu_generate_output_matches_reference(name):
    check_case(name, on_emulation)


@pytest.mark.gpu
@pytest.mark.parametrize('name', golden_names('io_'))
def test_gpu_generate_output_matches_reference(name):
    check_case(name, lambda dec: dec)


def check_drop_only_case(name, prepare):
    """decode(..., defects_possible=False) on a random third of the cycles (snowflake/main.py:244-264)."""
    meta, arr = load_golden(name)
    code = make_code(meta['spec'])
    dec = prepare(L.decoders.Snowflake(code, **meta['decoder_kwargs']))
    vectors, floors = strings(arr['vectors']), strings(arr['floors'])
    sets = dec._BITSTRING_CONVERTER.syndrome_vectors_to_sets(vectors)
    n = min(len(sets), 24)
    for mode in ('all', 'merging', 'unrooting'):
        dec.reset()
        dec.init_history()
        got = [dec.decode(s, time_only=mode, defects_possible=not d, log_floor_history=True)
               for s, d in zip(sets[:n], arr['drop_only'][:n])]
        assert got == arr[f'runtime_{mode}'][:n].tolist(), mode
        assert dec.floor_history == floors[:n]
    assert arr['drop_only'][:n].any() and not arr['drop_only'][:n].all()


@pytest.mark.parametrize('name', golden_names('iod_'))
def test_emu_decode_drop_only_cycles(name):
    check_drop_only_case(name, on_emulation)


@pytest.mark.gpu
@pytest.mark.parametrize('name', golden_names('iod_'))
def test_gpu_decode_drop_only_cycles(name):
    check_drop_only_case(name, lambda dec: dec)


def notebook_cases():
    rep = L.decoders.Snowflake(L.Repetition(5, 'phenomenological', scheme='frugal'), schedule='1:1')
    sf = L.decoders.Snowflake(L.Surface(3, 'circuit-level', scheme='frugal'), schedule='1:1')
    return [(rep, NOTEBOOK_REP5_IN, NOTEBOOK_REP5_OUT), (sf, NOTEBOOK_SF3_IN, NOTEBOOK_SF3_OUT)]


def test_emu_notebook_examples(tmp_path):
    for dec, vin, vout in notebook_cases():
        on_emulation(dec)
        path = tmp_path / 'snowflake_data.csv'
        assert dec.generate_output(list(vin), output_to_csv_file=str(path)) == vout
        rows = list(csv.reader(open(path)))
        h = dec.CODE.SCHEME.WINDOW_HEIGHT
        assert rows[0] == ['syndrome_vector', 'floor_history'] and len(rows) == 1 + len(vin) + h
        assert [r[0] for r in rows[1:]] == vin + h * ['0' * len(vin[0])]
        assert [r[1] for r in rows[1:]] == vout


@pytest.mark.gpu
def test_gpu_notebook_examples():
    for dec, vin, vout in notebook_cases():
        assert dec.generate_output(list(vin)) == vout


def test_bitstring_converter():
    rep = Bitstrings(5, 6, True)
    assert rep.LENGTH == 4 and rep.vector_to_set('1001') == {(0, 5), (3, 5)}
    assert rep.set_to_vector({(1, 5)}) == '0100'
    sf = Bitstrings(3, 4, False)
    assert sf.LENGTH == 6
    # west to east along a row, rows from south to north
    assert sf.vector_to_set('100000') == {(2, 0, 3)} and sf.vector_to_set('000001') == {(0, 1, 3)}
    for v in ('110010', '000000', '111111'):
        assert sf.set_to_vector(sf.vector_to_set(v)) == v
    with pytest.raises(ValueError):
        sf.vector_to_set('101')
    with pytest.raises(ValueError):
        sf.set_to_vector({(0, 0, 2)})          # not the top layer
    with pytest.raises(ValueError):
        rep.set_to_vector({(4, 5)})
    with pytest.raises(NotImplementedError):
        rep.lowest_edges('circuit-level')
    assert len(sf.lowest_edges('circuit-level')) == 28 and len(sf.lowest_edges('circuit-level', False)) == 22
    assert len(rep.lowest_edges('phenomenological')) == 9


def test_decode_argument_errors():
    dec = on_emulation(L.decoders.Snowflake(L.Surface(3, 'phenomenological', scheme='frugal')))
    with pytest.raises(ValueError):
        dec.decode({(0, 0, 0)})                # not in the top layer
    with pytest.raises(ValueError):
        dec.decode(set(), time_only='nothing')
    with pytest.raises(NotImplementedError):
        dec.decode(set(), log_history='fine')
    assert dec.decode(set(), time_only='all') == 4 and dec.decode(set()) == 0
