"""The kernel phase sources under AddressSanitizer + UBSan.

compute-sanitizer is not available on the GPU pool, so out-of-range indices of
the phase logic are hunted on the host: the emulation (same .cuh / .inl sources,
lanes run sequentially, one heap buffer per shot = the shared-memory slice) is
built with -fsanitize=address,undefined and re-runs a golden case of every
kernel family in a child process (ASan must be preloaded into python)."""
import os
import subprocess
import sys

import pytest

from conftest import ROOT

sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import emu_binding  # noqa: E402

CHILD = r'''
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(ROOT, 'tests'))
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
sys.path.insert(0, os.path.join(ROOT, 'oracle'))
from conftest import load_golden, make_code
import emu_binding, binding
for name, flags in (('uf_sf5_phen_p08', 0), ('uf_sf9_cc_p10', 1), ('uf_sf4_cl_fwd_p03', 2), ('nuf_sf5_cl_p02', 4),
                    ('nbuf_sf9_cc_p15', 12)):
    meta, arr = load_golden(name)
    emu = emu_binding.Emu(make_code(meta['spec']))
    v = {0: 'default', 1: 'dynamic', 2: 'west', 4: 'default', 12: 'default'}[flags]
    for decode in (emu.decode_uf, lambda f, s: emu.decode_uf_cta(f, s, nl=448)):
        corr, _, _ = decode(flags, arr['syn_bits'])
        assert np.array_equal(corr, arr[v + '_corr']), name
code = make_code(dict(code='Surface', d=5, noise='circuit-level'))
emu = emu_binding.Emu(code)
emu.mc_uf(0, 0.02, 5, 0, 200)
for nl in (8, 32, 128):                      # the compact fast path: fitted lists, starved lists (every overflow branch)
    emu.mc_uf_compact(0.02, 5, 0, 300, nl=nl)
    emu.mc_uf_compact(0.02, 5, 0, 300, nl=nl, mu=0.0)
emu_binding.Emu(make_code(dict(code='Surface', d=9, noise='phenomenological'))).mc_uf_compact(0.03, 5, 0, 200, nl=16)
emu.sample_weight(7, 3, 0, 100)
meta, arr = load_golden('ca_sf5_phen_p03')
emu = emu_binding.Emu(make_code(meta['spec']))
for key, (dec, fl) in (('macar', (1, 0)), ('actis', (2, 0))):
    if key in meta['decoders']:
        corr, _, steps = emu.decode_ca(dec, fl, arr['syn_bits'])
        assert np.array_equal(steps, arr[key + '_steps'])
meta, arr = load_golden('sf_sf5_cl')
es = emu_binding.EmuStream(make_code(meta['spec']))
fresh = binding.code_bits_to_local(es.tables, arr['fresh_bits'], 0)
rt, floor, m = es.stream(0, fresh)
assert np.array_equal(m, arr['m_cycle'])
meta, arr = load_golden('fw_sf5_phen')
code = make_code(meta['spec'])
emu = emu_binding.Emu(code)
steps, commit, m = emu.forward_stream(code, 0, 0, arr['uf_init'], arr['uf_fresh'])
assert np.array_equal(m, arr['uf_m'])
print('sanitized emulation ok')
'''


def test_phase_sources_under_asan_ubsan():
    lib, asan = emu_binding.build_sanitized()
    if not os.path.exists(asan):
        pytest.skip('libasan not found')
    env = dict(os.environ, LD_PRELOAD=asan, LUF_EMU_LIB=lib, ASAN_OPTIONS='detect_leaks=0:abort_on_error=0',
               UBSAN_OPTIONS='print_stacktrace=1')
    r = subprocess.run([sys.executable, '-c', f'ROOT = {ROOT!r}\n' + CHILD], env=env, capture_output=True, text=True,
                       cwd=ROOT, timeout=900)
    assert r.returncode == 0 and 'sanitized emulation ok' in r.stdout, (r.stdout[-2000:], r.stderr[-4000:])
