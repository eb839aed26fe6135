import sys, os
sys.path.insert(0, '.'); sys.path.insert(0, 'oracle'); sys.path.insert(0, 'tests')
import numpy as np
import binding
from conftest import make_code
from localuf_b200 import _engine
spec = dict(code='Surface', d=5, noise='circuit-level', kwargs=dict(scheme='forward')); p = 0.012; nstreams = 100; ncycles = 10
code = make_code(spec)
eng = _engine.Engine(code, device=0)
ft = eng.attach_forward()
orc = binding.ForwardOracle(code, ft)
g = code.FLAT
rng = np.random.default_rng(7)
probs = np.array(code.NOISE.class_probabilities(p))
def bits(mask):
    pad = np.zeros(mask.shape[:-1] + (32 * eng.WE,), np.uint8); pad[..., :g.n_edges] = mask
    return np.packbits(pad, axis=-1, bitorder='little').view(np.uint32)
fresh_p = np.where(g.edge_class != 255, probs[np.minimum(g.edge_class, len(probs) - 1)], 0.0)
buf_p = np.where(ft.buffer_class != 255, probs[np.minimum(ft.buffer_class, len(probs) - 1)], 0.0)
fresh = rng.random((nstreams, ncycles, g.n_edges)) < fresh_p
fresh[:, -ft.n_cleanse:, :] = False
init = rng.random((nstreams, g.n_edges)) < buf_p
fresh_b, init_b = bits(fresh), bits(init)
print('N', g.n_nodes, 'B', g.n_boundary, 'E', g.n_edges)
for dec in (0, 1):
    for rep in range(2):
        steps, commit, m = eng.forward_decode_host(dec, 0, init_b, fresh_b)
        bad = 0
        for k in range(nstreams):
            o_steps, o_commit, o_m = orc.stream(dec, 0, init_b[k], fresh_b[k])
            if not (np.array_equal(steps[k], o_steps) and np.array_equal(commit[k], o_commit) and np.array_equal(m[k], o_m)):
                bad += 1
                if bad <= 3:
                    cyc = [c for c in range(ncycles) if not np.array_equal(commit[k, c], o_commit[c]) or not np.array_equal(steps[k, c], o_steps[c])]
                    print('dec', dec, 'rep', rep, 'stream', k, 'first bad cycles', cyc[:4], 'steps gpu', steps[k, cyc[0]] if cyc else None, 'orc', o_steps[cyc[0]] if cyc else None)
        print('dec', dec, 'rep', rep, 'bad streams', bad)
    # single-stream launches
    bad1 = 0
    for k in range(20):
        steps, commit, m = eng.forward_decode_host(dec, 0, init_b[k:k+1], fresh_b[k:k+1])
        o_steps, o_commit, o_m = orc.stream(dec, 0, init_b[k], fresh_b[k])
        bad1 += not (np.array_equal(steps[0], o_steps) and np.array_equal(commit[0], o_commit))
    print('dec', dec, 'single-stream bad', bad1)
