#!/usr/bin/env python
"""BASELINE.json configs[4] in miniature: threshold sweep d x p through
sim.accuracy.monte_carlo (UF, phenomenological batch), timing every point.

    python tools/sweep_threshold.py [--ds 5,9,13,17,21,25] [--ps 0.015,0.026,0.035] [--seconds 1.0]
    python tools/sweep_threshold.py --decoder Snowflake --ps 0.004,0.008      # circuit-level, frugal scheme;
                                                                              # one shot = Frugal.run(n=1)
"""
import argparse
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import localuf_b200 as L


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--ds', default='5,9,13,17,21,25')
    ap.add_argument('--ps', default='0.015,0.026,0.035')
    ap.add_argument('--noise', default='phenomenological')
    ap.add_argument('--seconds', type=float, default=1.0)
    ap.add_argument('--decoder', default='UF', choices=['UF', 'Macar', 'Actis', 'Snowflake'])
    a = ap.parse_args()
    if a.decoder == 'Snowflake':
        return sweep_snowflake(a)
    print(f'{"d":>3s} {"p":>6s} {"N":>6s} {"E":>6s} {"shots":>9s} {"fails":>8s} {"rate":>9s} {"Mshots/s":>9s}')
    for d in (int(x) for x in a.ds.split(',')):
        code = L.Surface(d, a.noise)
        dec = getattr(L.decoders, a.decoder)(code)
        for p in (float(x) for x in a.ps.split(',')):
            n = 64 if a.decoder == 'Actis' else 2048
            code.SCHEME.run(dec, p, n, seed=1)               # warm up (graph upload, launch plan)
            t0 = time.perf_counter(); code.SCHEME.run(dec, p, n, seed=2); dt = time.perf_counter() - t0
            n = max(n, min(1 << 24, int(n * a.seconds / max(dt, 1e-4))))
            t0 = time.perf_counter(); m, _ = code.SCHEME.run(dec, p, n, seed=3); dt = time.perf_counter() - t0
            print(f'{d:3d} {p:6.3f} {code.FLAT.n_nodes:6d} {code.FLAT.n_edges:6d} {n:9d} {m:8d} {m / n:9.2e} {n / dt / 1e6:9.3f}',
                  flush=True)


def sweep_snowflake(a):
    print(f'{"d":>3s} {"p":>6s} {"h":>3s} {"nodes":>6s} {"runs":>8s} {"m":>7s} {"Mcycles/s":>10s}')
    for d in (int(x) for x in a.ds.split(',')):
        code = L.Surface(d, 'circuit-level', scheme='frugal')
        dec = L.decoders.Snowflake(code)
        h = code.SCHEME.WINDOW_HEIGHT
        cycles = h + d + d + 2 * h                            # of one Frugal.run(n=1)
        for p in (float(x) for x in a.ps.split(',')):
            n = 296
            code.SCHEME.run(dec, p, 1, shots=n, seed=1)
            t0 = time.perf_counter(); code.SCHEME.run(dec, p, 1, shots=n, seed=2); dt = time.perf_counter() - t0
            n = max(n, min(1 << 20, int(n * a.seconds / max(dt, 1e-4))))
            t0 = time.perf_counter(); m, _ = code.SCHEME.run(dec, p, 1, shots=n, seed=3); dt = time.perf_counter() - t0
            print(f'{d:3d} {p:6.3f} {h:3d} {dec._tables.NL * h:6d} {n:8d} {m:7d} {n * cycles / dt / 1e6:10.3f}', flush=True)


if __name__ == '__main__':
    main()
