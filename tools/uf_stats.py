#!/usr/bin/env python
"""Event counts of the UF kernel phases per shot, from the host emulation built
with -DLUF_STATS (test/analysis tooling; never a product path).

    python tools/uf_stats.py [workload] [shots]
"""
import ctypes as C
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, 'tests', 'emu'))
import numpy as np
import bench
import emu_binding

NAMES = {0: 'growth rounds', 1: 'active nodes, round 1', 2: 'active nodes, later rounds', 3: 'newly full edges',
         4: 'rounds with > 32 newly full edges', 5: 'reservation iterations', 6: '  of round 1', 7: '  of later rounds',
         8: 'newly full edges, round 1', 9: 'newly full edges, later rounds', 10: 'shots with a general peel',
         11: 'tree roots', 12: 'leaf edges', 13: 'folded stars'}


def main():
    wl = sys.argv[1] if len(sys.argv) > 1 else 'cfg3'
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 20000
    out = os.path.join(ROOT, 'tests', 'emu', '_build', 'libluf_emu_stats.so')
    subprocess.run(['g++', '-O2', '-fPIC', '-shared', '-std=c++17', '-DLUF_STATS', '-o', out,
                    os.path.join(ROOT, 'tests', 'emu', 'emu.cpp')], check=True)
    lib = C.CDLL(out)
    emu_binding._lib = lib
    code, p, desc = bench.make_code(wl)
    e = emu_binding.Emu(code)
    fails, _, _, _ = e.mc_uf(0, p, 1, 0, n, export=False)
    stats = (C.c_ulonglong * 32).in_dll(lib, 'luf_stats')
    print(desc, f'{n} shots, {fails} failures')
    for i, name in NAMES.items():
        if stats[i]:
            print(f'{name:40s} {stats[i] / n:10.3f} per shot')


if __name__ == '__main__':
    main()
