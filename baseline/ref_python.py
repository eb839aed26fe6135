#!/usr/bin/env python
"""The UNMODIFIED Python reference (timchan0/localuf) timed on the host cores.

`baseline/_ref/` (git-ignored, shipped to the GPU box by gpurun) holds the reference package exactly as
`pip install --target baseline/_ref /root/reference` would lay it out; __graft_entry__.build() puts it there
when /root/reference is present (the pip build itself needs `hatchling`, which is absent from the offline
wheelhouse -- DESIGN.md section 9 -- so the pure-Python package directory is copied as pip would).  Nothing of
this repository is on the path this file times: it is the reference's own `code.SCHEME.run(decoder, p, n)`
(`localuf/_schemes.py:49-76,116-121`), one independent process per host core, the reference having no
parallelism of its own.  Five optional dependencies of the reference that the image lacks (pymatching, stim,
matplotlib, statsmodels, ipywidgets: MWPM and plotting only) are stubbed.

    python baseline/ref_python.py --workload cfg3 --n 500 --seed 1        # one process, prints one JSON line
"""
from __future__ import annotations

import argparse
import json
import os
import random
import subprocess
import sys
import time
import types

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.path.join(HERE, '_ref')

# name: (code class, d, noise, scheme, decoder, p, units per run(n) call as f(n, d, h))
WORKLOADS = {
    'cfg1': ('Repetition', 5, 'code capacity', 'batch', 'UF', 0.05),
    'cfg2': ('Surface', 9, 'code capacity', 'batch', 'UF', 0.05),
    'cfg3': ('Surface', 11, 'phenomenological', 'batch', 'UF', 0.01),
    'cfg3-macar': ('Surface', 11, 'phenomenological', 'batch', 'Macar', 0.01),
    'cfg3-actis': ('Surface', 11, 'phenomenological', 'batch', 'Actis', 0.01),
    'cfg4': ('Surface', 13, 'circuit-level', 'frugal', 'Snowflake', 0.003),
}
DEFAULT_N = {'cfg1': 10000, 'cfg2': 20000, 'cfg3': 500, 'cfg3-macar': 100, 'cfg3-actis': 10, 'cfg4': 2}


def available() -> bool:
    return os.path.isdir(os.path.join(REF, 'localuf'))


def import_reference():
    class _Stub(types.ModuleType):
        def __getattr__(self, name):
            if name.startswith('__'):
                raise AttributeError(name)
            m = _Stub(self.__name__ + '.' + name)
            setattr(self, name, m)
            return m

        def __call__(self, *a, **k):
            raise RuntimeError('stubbed optional dependency of the reference')
    for name in ('pymatching', 'stim', 'matplotlib', 'matplotlib.pyplot', 'matplotlib.container', 'matplotlib.lines',
                 'statsmodels', 'statsmodels.api', 'statsmodels.stats', 'statsmodels.stats.proportion', 'ipywidgets'):
        try:
            __import__(name)
        except Exception:
            sys.modules[name] = _Stub(name)
    sys.path.insert(0, REF)
    import localuf
    assert os.path.realpath(localuf.__file__).startswith(os.path.realpath(REF)), localuf.__file__
    return localuf


def run_one(workload: str, n: int, seed: int, p=None, d=None) -> dict:
    localuf = import_reference()
    from localuf import decoders
    cls, d0, noise, scheme, dec, p0 = WORKLOADS[workload]
    d = d0 if d is None else d
    p = p0 if p is None else p
    random.seed(seed)
    code = getattr(localuf, cls)(d, noise, scheme=scheme)
    decoder = getattr(decoders, dec)(code)
    t0 = time.perf_counter()
    m, slenderness = code.SCHEME.run(decoder, p, n)
    dt = time.perf_counter() - t0
    if scheme == 'frugal':          # one run(n) = ceil(h/c) + n*d + d + 2h decoding cycles (_schemes.py:566-649)
        h = code.SCHEME.WINDOW_HEIGHT
        units = h + n * d + d + 2 * h
    else:
        units = n
    return {'units': units, 'failures': int(m), 'seconds': dt, 'n': n}


def time_parallel(workload: str, procs: int, n: int | None = None, p=None, d=None, timeout=600) -> dict:
    """`procs` independent processes, each code.SCHEME.run(decoder, p, n); aggregate units / the slowest
    process's run time (they start together)."""
    n = DEFAULT_N[workload] if n is None else n
    cmd = [sys.executable, os.path.abspath(__file__), '--workload', workload, '--n', str(n)]
    if p is not None:
        cmd += ['--p', str(p)]
    if d is not None:
        cmd += ['--d', str(d)]
    t0 = time.perf_counter()
    ps = [subprocess.Popen(cmd + ['--seed', str(1000 + k)], stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True)
          for k in range(procs)]
    rows = []
    for q in ps:
        out, err = q.communicate(timeout=timeout)
        if q.returncode != 0:
            raise RuntimeError('reference process failed: ' + err[-500:])
        rows.append(json.loads(out.strip().splitlines()[-1]))
    wall = time.perf_counter() - t0
    units = sum(r['units'] for r in rows)
    slowest = max(r['seconds'] for r in rows)
    return {'value': units / slowest, 'per_core': units / slowest / procs, 'cores': procs, 'units': units,
            'failures': sum(r['failures'] for r in rows), 'seconds': slowest, 'wall_seconds': wall, 'n_per_process': n}


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--n', type=int, default=None)
    ap.add_argument('--seed', type=int, default=1000)
    ap.add_argument('--p', type=float, default=None)
    ap.add_argument('--d', type=int, default=None)
    ap.add_argument('--procs', type=int, default=0, help='> 0: run that many processes and print the aggregate')
    a = ap.parse_args()
    if a.procs > 0:
        print(json.dumps(time_parallel(a.workload, a.procs, a.n, a.p, a.d)))
    else:
        print(json.dumps(run_one(a.workload, DEFAULT_N[a.workload] if a.n is None else a.n, a.seed, a.p, a.d)))
