import sys
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import binding
from conftest import make_code
from localuf_b200 import _engine
for spec in (dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='forward')),
             dict(code='Surface', d=5, noise='phenomenological', kwargs=dict(scheme='forward')),
             dict(code='Surface', d=5, noise='circuit-level')):
    code = make_code(spec)
    eng = _engine.Engine(code, device=0)
    orc = binding.Oracle.from_code(code)
    g = code.FLAT
    rng = np.random.default_rng(3)
    n = 20000
    flips = rng.random((n, g.n_edges)) < 0.02
    pad = np.zeros((n, 32 * eng.WE), np.uint8); pad[:, :g.n_edges] = flips
    err = np.packbits(pad, axis=1, bitorder='little').view(np.uint32)
    syn = orc.syndrome(err)
    for flags in (0, 1, 2):
        corr, growth2b, steps = eng.decode_batch_host(0, flags, syn, want_growth=True)
        ocorr, ogrowth, osteps = orc.decode_batch(0, flags, syn)
        gr = _engine.unpack_growth(growth2b, g.n_edges)
        badc = np.flatnonzero((corr != ocorr).any(axis=1)); badg = np.flatnonzero((gr != ogrowth).any(axis=1))
        print(spec['noise'], spec.get('kwargs'), 'flags', flags, 'bad corr', len(badc), 'bad growth', len(badg), 'bad rounds', int((steps[:, 0] != osteps[:, 0]).sum()), badc[:5])
        if len(badc) and flags == 0:
            k = badc[0]
            a = set(np.flatnonzero(np.unpackbits(corr[k].view(np.uint8), bitorder='little')[:g.n_edges]))
            b = set(np.flatnonzero(np.unpackbits(ocorr[k].view(np.uint8), bitorder='little')[:g.n_edges]))
            print('  shot', k, 'gpu-only', sorted(a - b), 'oracle-only', sorted(b - a), 'defects', int(np.unpackbits(syn[k].view(np.uint8)).sum()))
            for e in sorted((a - b) | (b - a)):
                print('   edge', e, g.edges[e])
