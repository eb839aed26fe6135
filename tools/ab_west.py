"""West-inclination fused Monte Carlo (the compact fast kernel alone) at a few sizes: A/B of two builds of the
library (LUF_B200_LIB)."""
import os, sys, time
sys.path.insert(0, os.getcwd())
import ctypes
_orig = ctypes.CDLL.__getattr__
class _Missing:            # an older build of the library lacks the newest entry points
    pass
def _tolerant(self, name):
    try:
        return _orig(self, name)
    except AttributeError:
        if name.startswith('luf_'):
            m = _Missing(); setattr(self, name, m); return m
        raise
ctypes.CDLL.__getattr__ = _tolerant
import localuf_b200 as L
from localuf_b200 import _engine
for d, n, p in ((17, 65536, 0.026), (17, 262144, 0.015), (13, 262144, 0.026), (11, 4194304, 0.01)):
    code = L.Surface(d, "phenomenological"); eng = _engine.Engine(code, device=0)
    eng.mc_run_host(0, 2, p, 1, 0, n)
    best = 0
    for k in range(3):
        t0 = time.perf_counter(); r = eng.mc_run_host(0, 2, p, 2, k * n, n); best = max(best, n / (time.perf_counter() - t0) / 1e6)
    print("RES d=%d p=%g west: %.2f M shots/s" % (d, p, best), r)
