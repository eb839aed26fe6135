import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, 'tests', 'golden')


def pytest_configure(config):
    config.addinivalue_line('markers', 'gpu: needs a CUDA device (run on the B200 box)')


def load_golden(name):
    """Return (meta dict, arrays) of ``tests/golden/<name>.npz``."""
    z = np.load(os.path.join(GOLDEN, name + '.npz'))
    meta = json.loads(str(z['meta']))
    return meta, {k: z[k] for k in z.files if k != 'meta'}


def golden_names(prefix):
    return sorted(f[:-4] for f in os.listdir(GOLDEN) if f.startswith(prefix) and f.endswith('.npz'))


def make_code(spec):
    import localuf_b200 as L
    cls = {'Surface': L.Surface, 'Repetition': L.Repetition}[spec['code']]
    return cls(spec['d'], spec['noise'], **spec.get('kwargs', {}))


@pytest.fixture(scope='session')
def golden():
    return load_golden
