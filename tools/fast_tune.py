#!/usr/bin/env python
"""Launch-shape sweep of the compact fast path (k_mc_uf_fast): tile width, adjacency placement, blocks per SM,
shots per block.  The plan's knobs are environment variables read once per process, so every variant runs in a
child process; the child times luf_mc_run_host over a few large batches (wall clock around a synchronous call,
no profiler).

    python tools/fast_tune.py --d 11 --p 0.01 [--noise phenomenological] [--tw 8,16,32] [--adj 0,1] [--bps 1,2]
"""
import argparse
import itertools
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def child(a):
    import localuf_b200 as L
    from localuf_b200 import _engine
    code = L.Surface(a.d, a.noise)
    eng = _engine.Engine(code, device=0)
    n = a.shots
    eng.mc_run_host(0, a.flags, a.p, 1, 0, max(1024, n // 8))
    best, fails = 0.0, 0
    for k in range(a.reps):
        t0 = time.perf_counter()
        fails, shots = eng.mc_run_host(0, a.flags, a.p, 2, k * n, n)
        dt = time.perf_counter() - t0
        best = max(best, n / dt)
    print(f'RESULT {best / 1e6:.3f} {fails}')


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--d', type=int, default=11)
    ap.add_argument('--p', type=float, default=0.01)
    ap.add_argument('--noise', default='phenomenological')
    ap.add_argument('--flags', type=int, default=0)
    ap.add_argument('--shots', type=int, default=1 << 22)
    ap.add_argument('--reps', type=int, default=3)
    ap.add_argument('--tw', default='8,16,32')
    ap.add_argument('--adj', default='0,1')
    ap.add_argument('--bps', default='1,2')
    ap.add_argument('--spb', default='0')
    ap.add_argument('--warps', default='')
    ap.add_argument('--baseline', action='store_true', help='also time LUF_FAST=0 (the one-launch kernel)')
    ap.add_argument('--child', action='store_true')
    a = ap.parse_args()
    if a.child:
        return child(a)
    base = [sys.executable, os.path.abspath(__file__), '--child', '--d', str(a.d), '--p', str(a.p), '--noise', a.noise,
            '--flags', str(a.flags), '--shots', str(a.shots), '--reps', str(a.reps)]

    def run(env_extra, label):
        env = dict(os.environ, LUF_DEBUG_PLAN='1', **env_extra)
        r = subprocess.run(base, env=env, capture_output=True, text=True)
        res = [l for l in r.stdout.splitlines() if l.startswith('RESULT')]
        plan = [l for l in r.stderr.splitlines() if 'fast plan' in l]
        if not res:
            print(f'{label:40s} FAILED {r.stderr[-300:]}', flush=True)
            return
        _, rate, fails = res[0].split()
        print(f'{label:40s} {float(rate):9.2f} M shots/s  fails {fails}  {plan[0].split("fast plan:")[1].strip() if plan else ""}', flush=True)

    print(f'# Surface d={a.d} {a.noise} p={a.p} flags={a.flags}, {a.shots} shots per call')
    if a.baseline:
        run({'LUF_FAST': '0'}, 'round-1 kernel (LUF_FAST=0)')
    run({}, 'default plan')
    if a.warps:
        for w in a.warps.split(','):
            run({'LUF_FAST_WARPS': w}, f'warps target {w}')
    for tw, adj, bps, spb in itertools.product(a.tw.split(','), a.adj.split(','), a.bps.split(','), a.spb.split(',')):
        if not tw:
            continue
        env = {'LUF_FAST_TW': tw, 'LUF_FAST_ADJ': adj, 'LUF_FAST_BPS': bps}
        if spb != '0':
            env['LUF_FAST_SPB'] = spb
        run(env, f'tw={tw} adj={adj} bps={bps} spb={spb}')


if __name__ == '__main__':
    main()
