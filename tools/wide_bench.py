#!/usr/bin/env python
"""Throughput of the wide instance of the UF kernels (luf_uf_wide.cuh) on graphs beyond the 16-bit layout:
the fused Monte-Carlo kernel (sample + canonical decode + logical bit) and the staged decode on device buffers,
for each placement of the state slab and block size.  The plan's knobs are read per launch, so one process
runs all variants; a launch is timed with CUDA events on torch's current stream.

    python tools/wide_bench.py --d 23 --p 0.004 [--noise circuit-level] [--shots 8192] [--flags 1]
"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--d', type=int, default=23)
    ap.add_argument('--p', type=float, default=0.004)
    ap.add_argument('--noise', default='circuit-level')
    ap.add_argument('--flags', type=int, default=1, help='UF flags of the fused run (1 = dynamic: the wide kernel alone)')
    ap.add_argument('--shots', type=int, default=8192)
    ap.add_argument('--force', action='store_true', help='LUF_FORCE_WIDE=1: a graph that also fits the 16-bit layout')
    ap.add_argument('--variants', default='global:256,global:512,smem:256,smem:512')
    a = ap.parse_args()
    if a.force:
        os.environ['LUF_FORCE_WIDE'] = '1'
    os.environ['LUF_DEBUG_PLAN'] = '1'
    import torch
    import localuf_b200 as L
    from localuf_b200 import _engine
    code = L.Surface(a.d, a.noise)
    eng = _engine.Engine(code, device=0)
    dev = torch.device('cuda', 0)
    n = a.shots
    print(f'# Surface d={a.d} {a.noise} p={a.p}: {code.N_EDGES} edges, {code.FLAT.n_nodes} nodes, {n} shots per launch')
    counts = torch.zeros(4, dtype=torch.int64, device=dev)
    err = torch.empty((n, eng.WE), dtype=torch.int32, device=dev)
    syn = torch.empty((n, eng.WD), dtype=torch.int32, device=dev)
    corr = torch.empty((n, eng.WE), dtype=torch.int32, device=dev)
    eng.sample(a.p, 1, 0, n, err, syn)

    def timed(f, reps=3):
        f()
        torch.cuda.synchronize()
        best = 1e30
        for _ in range(reps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); f(); e1.record()
            torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e-3)
        return best

    for var in a.variants.split(','):
        state, threads = var.split(':')
        os.environ['LUF_WIDE_STATE'] = state
        os.environ['LUF_WIDE_THREADS'] = threads
        try:
            counts.zero_()
            t_mc = timed(lambda: eng.mc_run(_engine.DEC_UF, a.flags, a.p, 2, 0, n, counts))
            t_dec = timed(lambda: eng.decode_batch(_engine.DEC_UF, 0, n, syn, corr))
            fails = torch.zeros(1, dtype=torch.int64, device=dev)
            eng.check(n, err, corr, None, fails)
            print(f'{var:14s} fused flags={a.flags}: {n / t_mc / 1e3:9.1f} k shots/s   staged decode: {n / t_dec / 1e3:9.1f} k shots/s   '
                  f'(failures of the staged run {int(fails.item())} / {n})', flush=True)
        except Exception as ex:          # a placement that does not fit
            print(f'{var:14s} {type(ex).__name__}: {ex}', flush=True)
    os.environ.pop('LUF_WIDE_STATE'); os.environ.pop('LUF_WIDE_THREADS')
    t_def = timed(lambda: eng.mc_run(_engine.DEC_UF, 0, a.p, 2, 0, n, counts))
    print(f'default plan, fused flags=0 (compact fast launch + wide launch over its HARD shots): {n / t_def / 1e3:9.1f} k shots/s')
    t_w = timed(lambda: eng.mc_run(_engine.DEC_UF, 2, a.p, 2, 0, n, counts))
    print(f'default plan, fused flags=2 (west: compact fast launch alone): {n / t_w / 1e3:9.1f} k shots/s')


if __name__ == '__main__':
    main()
