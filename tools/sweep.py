#!/usr/bin/env python
"""BASELINE.json configs[4]: the threshold sweep d in {5..25} x 8 noise levels, UF (phenomenological, batch
scheme) and Snowflake (circuit-level, frugal scheme), through the reference's own driver
``sim.accuracy.monte_carlo`` (localuf/sim/accuracy/monte_carlo.py:25-101).

    python tools/sweep.py [--decoder UF|Snowflake|both] [--ds 5,7,...] [--seconds 1.0] [--max-shots N]
                          [--resume gpurun_out/sweep_resume.json] [--out gpurun_out/sweep]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 \
           tools/sweep.py ...                      # the shots of every point shard over the ranks

UF: per distance, rounds of ONE ``monte_carlo({d: [(p, n_p), ...]}, Surface, UF, 'phenomenological')`` call over
the distance's open points -- a warm-up and a probe round of a few thousand shots, then the main round with n_p
chosen from the probe so that a point takes about ``--seconds`` (the 10^9 shots per point of BASELINE.json need
hours near threshold at d = 25; reduced counts are what this tool is for).  ``random.seed(seed + round)``
precedes every call -- the reference's reproducibility contract; under torchrun ``SCHEME.run`` shards the shots
of every point over the ranks and all-reduces the counters (NCCL).  The time of a point is the wall time of
its ``Scheme.run`` call inside ``monte_carlo`` (a timing wrapper around ``Batch.run`` / ``Frugal.run``; maximum
over the ranks).  Snowflake: ``monte_carlo``'s n is the LENGTH of one memory experiment (``Frugal.run``,
_schemes.py:566-649) -- one window, one thread block; the sweep runs many experiments of n = 1 side by side
instead (``SCHEME.run(decoder, p, 1, shots=n)``).  After every distance (UF) / point (Snowflake) rank 0 rewrites
the resume file ``{point: {seed, rounds, shots_done, fails, seconds}}``; a restarted sweep skips what is done
and continues the round numbering.  Output: a table per decoder with shots, failures, rate and units/s per
point, and the crossing of consecutive distances.
"""
import argparse
import json
import math
import os
import random
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

GRIDS = {
    'UF': dict(noise='phenomenological', scheme='batch', ps=[0.015 + k * (0.035 - 0.015) / 7 for k in range(8)]),
    'Snowflake': dict(noise='circuit-level', scheme='frugal', ps=[0.004 + k * (0.011 - 0.004) / 7 for k in range(8)]),
}


def crossing(ps, lo_rates, hi_rates):
    """Noise level where the curves of two distances cross (log-linear interpolation); None if they do not."""
    for k in range(len(ps) - 1):
        a0, a1, b0, b1 = lo_rates[k], lo_rates[k + 1], hi_rates[k], hi_rates[k + 1]
        if min(a0, a1, b0, b1) <= 0:
            continue
        f0, f1 = math.log(b0 / a0), math.log(b1 / a1)          # < 0 below threshold (larger d is better), > 0 above
        if f0 <= 0 < f1:
            return ps[k] + (ps[k + 1] - ps[k]) * (-f0) / (f1 - f0)
    return None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--decoder', default='both', choices=['UF', 'Snowflake', 'both'])
    ap.add_argument('--ds', default='5,7,9,11,13,15,17,19,21,23,25')
    ap.add_argument('--seconds', type=float, default=1.0, help='time budget per point (after the probe round)')
    ap.add_argument('--max-shots', type=int, default=1 << 26)
    ap.add_argument('--probe', type=int, default=0, help='shots of the probe round (0 = by decoder)')
    ap.add_argument('--seed', type=int, default=20260101)
    ap.add_argument('--resume', default=os.path.join(ROOT, 'gpurun_out', 'sweep_resume.json'))
    ap.add_argument('--out', default=os.path.join(ROOT, 'gpurun_out', 'sweep'))
    a = ap.parse_args()

    import torch
    import torch.distributed as dist
    import localuf_b200 as L
    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    state = {}
    if os.path.exists(a.resume):
        with open(a.resume) as f:
            state = json.load(f)

    def save():
        if rank == 0:
            os.makedirs(os.path.dirname(a.resume), exist_ok=True)
            tmp = a.resume + '.tmp'
            with open(tmp, 'w') as f:
                json.dump(state, f, indent=1)
            os.replace(tmp, a.resume)

    # per-call wall time of Scheme.run, taken inside monte_carlo (it builds code, decoder and engine once per
    # distance and then calls run for every (p, n) of that distance)
    run_times = []

    def timed(cls):
        inner = cls.run

        def run(self, *args, **kw):
            getattr(args[0], 'engine', None)            # the decoder's engine (graph upload) outside the timed call
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            out = inner(self, *args, **kw)
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            run_times.append(float(dt.item()))
            return out
        cls.run = run
    timed(L.schemes.Batch)
    timed(L.schemes.Frugal)

    def book(entry, n, m, dt):
        entry['rounds'] += 1; entry['shots_done'] += n; entry['fails'] += int(m); entry['seconds'] += dt

    def report(dec_name, d, p, entry):
        if rank == 0:
            timed_shots = entry['shots_done'] - entry.get('probe_shots', 0)
            print(f"{dec_name:9s} d={d:2d} p={p:.5f} shots={entry['shots_done']:10d} fails={entry['fails']:9d} "
                  f"rate={entry['fails'] / entry['shots_done']:.3e} {timed_shots / max(entry['seconds'], 1e-9) / 1e6:9.4f} M shots/s",
                  flush=True)

    names = ['UF', 'Snowflake'] if a.decoder == 'both' else [a.decoder]
    ds = [int(x) for x in a.ds.split(',')]
    for dec_name in names:
        grid = GRIDS[dec_name]
        ps = [round(p, 6) for p in grid['ps']]
        for d in ds:
            entries = {p: state.setdefault(f'{dec_name}|{d}|{p}', {'seed': a.seed + 7919 * d + int(round(p * 1e6)), 'rounds': 0,
                                                                  'shots_done': 0, 'fails': 0, 'seconds': 0.0, 'target_seconds': 0.0})
                       for p in ps}
            todo = [p for p in ps if not (entries[p]['rounds'] >= 2 and entries[p]['target_seconds'] >= a.seconds)]
            if not todo:
                continue                                         # done in an earlier run
            probe = a.probe or (256 * world if dec_name == 'Snowflake' else 16384 * world)
            if dec_name == 'UF':
                # rounds of ONE monte_carlo call over the distance's open points (warm-up, probe, main)
                def mc_round(counts):
                    """counts: {p: n}; every point uses its own round number as the seed offset"""
                    run_times.clear()
                    random.seed(entries[todo[0]]['seed'] + entries[todo[0]]['rounds'])
                    df = L.sim.accuracy.monte_carlo({d: [(p, counts[p]) for p in todo]}, L.Surface, L.decoders.UF,
                                                    grid['noise'], scheme=grid['scheme'])
                    return {p: (int(df.loc['m', (d, p)]), run_times[k]) for k, p in enumerate(todo)}
                fresh = [p for p in todo if entries[p]['rounds'] == 0]
                if fresh:
                    out = mc_round({p: probe for p in todo})                 # warm-up: graph upload, launch plans
                    for p in fresh:
                        book(entries[p], probe, out[p][0], 0.0)
                        entries[p]['probe_shots'] = probe
                out = mc_round({p: probe for p in todo})                     # probe: the time of a small call
                n_of = {}
                for p in todo:
                    n_of[p] = int(min(a.max_shots, max(probe, probe * a.seconds / max(out[p][1], 1e-5))))
                    n_of[p] -= n_of[p] % world
                    book(entries[p], probe, out[p][0], 0.0)
                    entries[p]['probe_shots'] = entries[p].get('probe_shots', 0) + probe
                out = mc_round(n_of)
                for p in todo:
                    book(entries[p], n_of[p], out[p][0], out[p][1])
                    entries[p]['target_seconds'] = a.seconds
                    report(dec_name, d, p, entries[p])
                save()
            else:
                # Snowflake: monte_carlo's n is the LENGTH of one memory experiment (Frugal.run, _schemes.py:566-649),
                # one window = one thread block; the sweep runs many experiments of n = 1 side by side instead
                code = L.Surface(d, grid['noise'], scheme='frugal')
                decoder = L.decoders.Snowflake(code)
                for p in todo:
                    entry = entries[p]

                    def sf_round(n):
                        run_times.clear()
                        random.seed(entry['seed'] + entry['rounds'])
                        m, _ = code.SCHEME.run(decoder, p, 1, shots=n)
                        return int(m), run_times[-1]
                    if entry['rounds'] == 0:
                        m, _ = sf_round(probe)
                        book(entry, probe, m, 0.0)
                        entry['probe_shots'] = probe
                    m, t = sf_round(probe)
                    book(entry, probe, m, 0.0)
                    entry['probe_shots'] = entry.get('probe_shots', 0) + probe
                    n = int(min(a.max_shots, max(probe, probe * a.seconds / max(t, 1e-5))))
                    n -= n % world
                    m, t = sf_round(max(n, world))
                    book(entry, max(n, world), m, t)
                    entry['target_seconds'] = a.seconds
                    report(dec_name, d, p, entry)
                    save()
    if rank == 0:
        for dec_name in names:
            grid = GRIDS[dec_name]
            ps = [round(p, 6) for p in grid['ps']]
            unit = 'memory experiments (Frugal.run(n=1))' if dec_name == 'Snowflake' else 'shots'
            lines = [f'# {dec_name}, Surface, {grid["noise"]}, scheme {grid["scheme"]}; {world} GPU(s); unit = {unit}',
                     '# d  p  shots  fails  rate  M_units_per_s']
            rates = {}
            for d in ds:
                for p in ps:
                    e = state.get(f'{dec_name}|{d}|{p}')
                    if not e:
                        continue
                    timed = e['shots_done'] - e.get('probe_shots', 0)
                    rates.setdefault(d, []).append(e['fails'] / e['shots_done'])
                    lines.append(f"{d:2d} {p:.5f} {e['shots_done']:11d} {e['fails']:10d} {e['fails'] / e['shots_done']:.4e} "
                                 f"{timed / max(e['seconds'], 1e-9) / 1e6:.4f}")
            xs = []
            for lo, hi in zip(ds, ds[1:]):
                if len(rates.get(lo, [])) == len(ps) == len(rates.get(hi, [])):
                    x = crossing(ps, rates[lo], rates[hi])
                    lines.append(f'# crossing d={lo}/{hi}: {x if x is None else round(x, 5)}')
                    if x:
                        xs.append(x)
            if xs:
                xs.sort()
                lines.append(f'# median crossing: {xs[len(xs) // 2]:.5f}')
            with open(f'{a.out}_{dec_name}.txt', 'w') as f:
                f.write('\n'.join(lines) + '\n')
            print('\n'.join(lines[-14:]))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
