"""Import shim for the UNMODIFIED upstream reference (timchan0/localuf).

TEST INFRASTRUCTURE ONLY.  Nothing in the product package imports this file.
It is used (a) by ``oracle/gen_golden.py`` in the authoring container to export
golden vectors from the real reference into ``tests/golden/`` and (b) by the
``not gpu`` tests that cross-check the host graph layer when ``/root/reference``
happens to be present (they skip otherwise -- the GPU box has no reference).

The reference needs five optional packages that are absent offline
(pymatching, stim, matplotlib, statsmodels, ipywidgets); none is on the
hot path, so empty stub modules are registered before the import.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_ROOT = os.environ.get("LOCALUF_REFERENCE", "/root/reference")


class _Stub(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        child = _Stub(self.__name__ + "." + name)
        setattr(self, name, child)
        return child

    def __call__(self, *a, **k):
        raise RuntimeError("stubbed optional dependency of the reference")


_OPTIONAL = [
    "pymatching", "stim", "matplotlib", "matplotlib.pyplot", "matplotlib.container",
    "matplotlib.lines", "statsmodels", "statsmodels.api", "statsmodels.stats",
    "statsmodels.stats.proportion", "ipywidgets",
]


def available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "localuf"))


def load():
    """Return the imported reference package ``localuf`` (or raise ImportError)."""
    if not available():
        raise ImportError(f"reference not found under {REFERENCE_ROOT}")
    for name in _OPTIONAL:
        try:
            __import__(name)
        except Exception:
            sys.modules.setdefault(name, _Stub(name))
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import localuf  # noqa: E402
    return localuf
