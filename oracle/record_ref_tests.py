"""Export the KNOWN ANSWERS of the reference's own test suite as a fixture.
TEST INFRASTRUCTURE ONLY (authoring container; needs /root/reference).

The reference's integration tests assert exact timestep counts, erasures and
corrections for hand-built syndromes.  Instead of
re-typing them, this script runs those UNMODIFIED test files under pytest with
the decoders' entry points wrapped by recorders: every call a *passing* test
made to ``validate`` / ``decode`` on a freshly reset decoder is stored as
(code spec, decoder, syndrome) -> (timesteps, erasure, correction), together
with the id of the test that made it -- what the test asserted is what was
recorded.  Likewise ``Snowflake._LOWEST_EDGES`` of every Snowflake the tests
build and every ``LogicalCounter.count`` call.

    python oracle/record_ref_tests.py        ->  tests/golden/ref_known_answers.json
"""
from __future__ import annotations

import functools
import json
import os
import sys

os.environ['PYTHONDONTWRITEBYTECODE'] = '1'
sys.dont_write_bytecode = True
HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

localuf = ref_shim.load()
import pytest  # noqa: E402
from localuf import Repetition, Surface  # noqa: E402
from localuf._base_classes import Code  # noqa: E402
from localuf._pairs import LogicalCounter, Pairs  # noqa: E402
from localuf.constants import Growth  # noqa: E402
from localuf.decoders import BUF, NodeBUF, NodeUF, UF, Snowflake  # noqa: E402
from localuf.decoders.luf import Actis, Macar  # noqa: E402
from localuf.decoders.luf.main import LUF, UnoptimalWaiter  # noqa: E402

TEST_FILES = [
    'tests/decoders/luf/main/LUF/test_LUF_integration.py',
    'tests/decoders/luf/main/MacarNode/test_MacarNode_integration.py',
    'tests/decoders/_inclination',
    'tests/decoders/node_uf/test_NodeUF.py',
    'tests/decoders/node_uf/NodeBUF/test_NodeBUF_integration.py',
    'tests/decoders/uf/BUF/test_BUF_integration.py',
    'tests/decoders/uf/UF/test_UF_integration.py',
    'tests/decoders/uf/UF/test_UF_unit.py',
    'tests/decoders/snowflake/Snowflake/test_Snowflake_integration.py',
    'tests/_pairs/test_LogicalCounter.py',
]
OUT = os.path.join(os.path.dirname(HERE), 'tests', 'golden', 'ref_known_answers.json')

records: list = []
state = {'nodeid': None, 'depth': 0, 'pending': []}


def spec_of(code):
    """Constructor arguments that rebuild `code`, or None if it is not a plain rebuildable code."""
    if not isinstance(code, Code) or type(code) not in (Surface, Repetition):
        return None
    scheme = str(code.SCHEME)
    candidates = [dict(scheme=scheme)]
    if scheme in ('forward', 'frugal'):
        candidates.append(dict(scheme=scheme, commit_height=code.SCHEME._COMMIT_HEIGHT,
                               buffer_height=code.SCHEME._BUFFER_HEIGHT))
    else:
        candidates.append(dict(scheme=scheme, window_height=code.SCHEME.WINDOW_HEIGHT))
    for kwargs in candidates:
        try:
            again = type(code)(code.D, str(code.NOISE), **kwargs)
        except Exception:
            continue
        if again.EDGES == code.EDGES and again.NODES == code.NODES:
            return dict(code=type(code).__name__, d=code.D, noise=str(code.NOISE), kwargs=kwargs)
    return None


def is_syndrome(s):
    return isinstance(s, (set, frozenset)) and all(isinstance(v, tuple) and all(isinstance(x, int) for x in v) for v in s)


def fresh(dec):
    try:
        return all(g is Growth.UNGROWN for g in dec.growth.values()) and not dec.syndrome
    except Exception:
        return False


def decoder_tag(dec):
    name = type(dec).__name__
    kw = {}
    if isinstance(dec, UF):
        kw = dict(dynamic='Dynamic' in type(dec._FORESTER).__name__,
                  inclination='west' if 'West' in type(dec._INCLINATION).__name__ else 'default')
    if isinstance(dec, Actis):
        kw = dict(_optimal=not isinstance(dec.NODES.WAITER, UnoptimalWaiter))
    return name, kw


def edges(es):
    return sorted([list(map(list, e)) for e in es])


def wrap(cls, method, kind):
    orig = cls.__dict__[method]

    @functools.wraps(orig)
    def wrapper(self, syndrome, *a, **k):
        outer = state['depth'] == 0
        ok = outer and type(self) in (UF, NodeUF, BUF, NodeBUF, Macar, Actis) and is_syndrome(syndrome) and fresh(self)
        spec = spec_of(self.CODE) if ok else None
        state['depth'] += 1
        try:
            out = orig(self, syndrome, *a, **k)
        finally:
            state['depth'] -= 1
        if spec is not None:
            try:
                name, kw = decoder_tag(self)
            except Exception:
                return out
            rec = dict(kind=kind, test=state['nodeid'], spec=spec, decoder=name, decoder_kwargs=kw,
                       syndrome=sorted(map(list, syndrome)),
                       erasure=edges(e for e, g in self.growth.items() if g is Growth.FULL),
                       full_or_burnt=edges(e for e, g in self.growth.items() if g in (Growth.FULL, Growth.BURNT)))
            if isinstance(self, LUF):
                rec['steps'] = list(out) if isinstance(out, tuple) else [int(out)]
            if kind == 'decode':
                rec['correction'] = edges(self.correction)
            state['pending'].append(rec)
        return out
    setattr(cls, method, wrapper)


for cls in (UF, BUF, NodeBUF, LUF):
    for method, kind in (('validate', 'validate'), ('decode', 'decode')):
        if method in cls.__dict__:
            wrap(cls, method, kind)

_sf_init = Snowflake.__init__


@functools.wraps(_sf_init)
def sf_init(self, code, *a, **k):
    _sf_init(self, code, *a, **k)
    spec = spec_of(code)
    if spec is not None and not a and set(k) <= {'merger', 'schedule', 'unrooter', '_include_timelike_lowest_edges'}:
        state['pending'].append(dict(kind='lowest_edges', test=state['nodeid'], spec=spec, decoder='Snowflake',
                                     decoder_kwargs=dict(k),
                                     lowest_edges=[list(map(list, e.INDEX)) for e in self._LOWEST_EDGES]))


Snowflake.__init__ = sf_init

_count = LogicalCounter.count


@functools.wraps(_count)
def count(self, pairs):
    before = sorted([list(u), list(v)] for u, v in pairs._dc.items()) if isinstance(pairs, Pairs) else None
    out = _count(self, pairs)
    if before is not None and isinstance(out, tuple) and isinstance(out[1], Pairs):
        state['pending'].append(dict(
            kind='logical_count', test=state['nodeid'],
            counter=dict(d=self._D, commit_height=self._COMMIT_HEIGHT, long_axis=self._LONG_AXIS, time_axis=self._TIME_AXIS),
            pairs=before, count=int(out[0]), new_pairs=sorted([list(u), list(v)] for u, v in out[1]._dc.items())))
    return out


LogicalCounter.count = count


class Recorder:
    def pytest_runtest_setup(self, item):
        state['nodeid'] = item.nodeid
        state['pending'] = []

    def pytest_runtest_logreport(self, report):
        if report.when == 'call' and report.passed:      # only what a passing test saw
            records.extend(state['pending'])
        if report.when == 'call':
            state['pending'] = []


def main():
    os.chdir(ref_shim.REFERENCE_ROOT)
    rc = pytest.main(['-q', '-p', 'no:cacheprovider', '--no-header', *TEST_FILES], plugins=[Recorder()])
    seen, unique = set(), []
    for r in records:
        key = json.dumps({k: v for k, v in r.items() if k != 'test'}, sort_keys=True)
        if key not in seen:
            seen.add(key)
            unique.append(r)
    with open(OUT, 'w') as f:
        json.dump(dict(source='recorded from the unmodified reference test suite by oracle/record_ref_tests.py',
                       pytest_exit_code=int(rc), records=unique), f, separators=(',', ':'))
    kinds = {}
    for r in unique:
        kinds[r['kind'] + ':' + r.get('decoder', '')] = kinds.get(r['kind'] + ':' + r.get('decoder', ''), 0) + 1
    print(f'{len(unique)} records ({len(records)} calls) -> {OUT}: {kinds}; {os.path.getsize(OUT) / 1024:.1f} KiB')


if __name__ == '__main__':
    main()
