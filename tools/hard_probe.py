#!/usr/bin/env python
"""One fused Monte-Carlo call near threshold (default inclination), for a launch list under ncu:
    ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/hard_probe.py 17 65536"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import localuf_b200 as L
from localuf_b200 import _engine

d, n = int(sys.argv[1]), int(sys.argv[2])
p = float(sys.argv[3]) if len(sys.argv) > 3 else 0.026
code = L.Surface(d, 'phenomenological')
eng = _engine.Engine(code, device=0)
print(eng.mc_run_host(0, 0, p, 2, 0, n))
