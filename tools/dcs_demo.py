#!/usr/bin/env python
"""Sample, decode and score shots on the device (UF.sample_confidence_scores): the command the k_dcs profile is
taken with.   python tools/dcs_demo.py [d] [p] [n]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import localuf_b200 as L


def main():
    d = int(sys.argv[1]) if len(sys.argv) > 1 else 11
    p = float(sys.argv[2]) if len(sys.argv) > 2 else 0.01
    n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 18
    code = L.Surface(d, 'phenomenological')
    dec = L.decoders.UF(code)
    dec.sample_confidence_scores(p, 4096, noise_level_for_priors=p, seed=1)
    t0 = time.perf_counter()
    out = dec.sample_confidence_scores(p, n, noise_level_for_priors=p, seed=2)
    dt = time.perf_counter() - t0
    print(f'd={d} p={p}: {n / dt / 1e6:.3f} M scored shots/s; mean swim distance {out["swim_distance"].mean():.3f}, '
          f'unclustered edge fraction {out["unclustered_edge_fraction"].mean():.4f}, '
          f'failure rate {out["logical"].mean():.2e}')


if __name__ == '__main__':
    main()
