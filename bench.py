#!/usr/bin/env python
"""Benchmark of the Monte-Carlo decode hot path (BASELINE.json metric:
decoded shots/s vs the reference on CPU, and fraction of the HBM roofline).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
                    [--workload cfg3|cfg2|cfg1] [--shots-per-step S]

A *step* is one pass of the fused hot path (sample noise -> syndrome -> UF
decode -> logical-failure count) over S synthetic shots per GPU.  The default
workload is BASELINE.json's headline configuration (configs[2], the one its
target is quoted on): Surface d=11, phenomenological noise p=0.01, batch scheme
(h=11), UF decoder.  Prints ONE JSON line on rank 0.

Under torchrun (N>1) every rank runs its own shot range of the global
counter-based RNG stream (weak scaling: S shots per GPU per step); the only
collective is the all-reduce of the (failures, shots) counters.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (code class, d, noise, p, description)
    'cfg1': ('Repetition', 5, 'code capacity', 0.05, 'Repetition d=5 code-capacity p=0.05 UF batch'),
    'cfg2': ('Surface', 9, 'code capacity', 0.05, 'Surface d=9 code-capacity p=0.05 UF batch'),
    'cfg3': ('Surface', 11, 'phenomenological', 0.01, 'Surface d=11 phenomenological p=0.01 h=11 UF batch'),
}
SEED = 20260101


def make_code(name):
    import localuf_b200 as L
    cls, d, noise, p, desc = WORKLOADS[name]
    return getattr(L, cls)(d, noise), p, desc


def algorithmic_bytes_per_shot(flat) -> int:
    """B = 4 W(E) + 2 W(D) bytes (error written and re-read,
    correction written and read, defects written and read), W(x) = 4 ceil(x/32)."""
    we = 4 * ((flat.n_edges + 31) // 32)
    wd = 4 * ((flat.n_detectors + 31) // 32)
    return 4 * we + 2 * wd


def measured_peak_hbm():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ['nvidia-smi', f'--id={self.index}', f'--query-gpu={self.Q}', '--format=csv,noheader,nounits', '-lms', '100'],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], [], set()
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        for r in self.rows:
            parts = [x.strip() for x in r.split(',')]
            if len(parts) < 6:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for n, v in zip(names, parts[2:6]):
                if v.lower().startswith('active'):
                    reasons.add(n)
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(mx) if mx else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


def cpu_oracle(code):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import binding
    return binding.Oracle.from_code(code)


def time_cpu(orc, class_p, target_s, threads, seed=SEED):
    """shots/s of the oracle port on `threads` host threads over ~target_s seconds."""
    n = 2000 * threads
    t0 = time.perf_counter(); orc.mc_batch(0, 0, class_p, seed, n, threads); dt = time.perf_counter() - t0
    n = max(n, int(n * target_s / max(dt, 1e-3)))
    t0 = time.perf_counter(); m = orc.mc_batch(0, 0, class_p, seed + 1, n, threads); dt = time.perf_counter() - t0
    return n / dt, n, m, dt


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The
    reference is pure Python and cannot travel to the GPU box, so this is the
    C restatement in oracle/ (pinned to reference goldens), kind = "port"."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    code, p, desc = make_code(args.workload)
    orc = cpu_oracle(code)
    threads = len(os.sched_getaffinity(0))
    class_p = code.NOISE.class_probabilities(p)
    for _ in range(args.warmup):
        orc.mc_batch(0, 0, class_p, 1, 500 * threads, threads)
    per_step_target = max(2.0, min(20.0, 60.0 / max(args.steps, 1)))
    rate0, n, _, _ = time_cpu(orc, class_p, per_step_target, threads)
    times = []
    for k in range(args.steps):
        t0 = time.perf_counter(); orc.mc_batch(0, 0, class_p, SEED + 10 + k, n, threads)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = n * args.steps / total
    print(json.dumps({
        'impl': 'reference', 'metric': 'decoded shots/s', 'value': value, 'unit': 'shots/s',
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'u32 bit-packed / u16 indices', 'data': 'synthetic',
        'config': {'workload': desc, 'shots_per_step': n, 'rng': 'xoshiro256** per thread'},
        'cpu_baseline': {'value': value, 'unit': 'shots/s', 'cores': threads, 'kind': 'port',
                         'sample': f'{n} shots per step on {threads} threads (C oracle of the reference algorithm)'},
        'e2e': {'value': value, 'unit': 'shots/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


def run_ours(args):
    import torch
    import torch.distributed as dist
    import localuf_b200 as L
    from localuf_b200 import _engine

    world = int(os.environ.get('WORLD_SIZE', 1))
    rank = int(os.environ.get('RANK', 0))
    local_rank = int(os.environ.get('LOCAL_RANK', 0))
    if not torch.cuda.is_available():
        raise SystemExit('bench.py needs a CUDA device: the engine has no CPU fallback')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))

    code, p, desc = make_code(args.workload)
    decoder = L.decoders.UF(code)
    eng = _engine.Engine(code, device=local_rank)
    decoder._engine = eng
    flat = code.FLAT
    S = args.shots_per_step
    counts = torch.zeros(4, dtype=torch.int64, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')     # > 126 MB L2

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def step(k):
        # rank r, step k owns global shots [(k*world + r)*S, +S)
        eng.mc_run(_engine.DEC_UF, 0, p, SEED, (k * world + rank) * S, S, counts)

    for k in range(args.warmup):
        step(k)
    barrier()
    counts.zero_()
    launches0 = eng.launch_count
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        t_wall0 = time.perf_counter()
        for k in range(args.steps):
            flush.fill_(k & 0xFF)                       # L2 flush between timed steps
            ev[k][0].record()
            step(args.warmup + k)
            ev[k][1].record()
        barrier()
        t_wall = time.perf_counter() - t_wall0
    launches = eng.launch_count - launches0
    step_ms = [a.elapsed_time(b) for a, b in ev]
    total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_ms = float(total_ms.item())
    fails, shots = (int(x) for x in counts.tolist()[:2])
    assert shots == S * args.steps * world, (shots, S, args.steps, world)
    value = shots / (total_ms / 1e3)

    # ---- end to end through the public API (the call a reference user makes)
    e2e_n = S
    for _ in range(2):
        code.SCHEME.run(decoder, p, min(e2e_n, 1 << 16), seed=1)
    barrier()
    t0 = time.perf_counter()
    e2e_steps = max(1, min(args.steps, 3))
    m_e2e = 0
    for k in range(e2e_steps):
        m, _ = code.SCHEME.run(decoder, p, e2e_n * world, seed=SEED + 1000 + k)
        m_e2e += m
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = e2e_n * world * e2e_steps / float(t_e2e.item())

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        bps = algorithmic_bytes_per_shot(flat)
        kern_ms = statistics.mean(step_ms)
        achieved = bps * S / (kern_ms / 1e3) / 1e9
        line = {
            'metric': 'decoded shots/s', 'value': value, 'unit': 'shots/s', 'n_gpus': world,
            'steps': args.steps, 'warmup': args.warmup, 'ms_per_step': total_ms / args.steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': 'u32 bit-packed / u16 indices', 'data': 'synthetic',
            'config': {'workload': desc, 'shots_per_step_per_gpu': S, 'seed': SEED,
                       'parallelism': f'dp{world} (shots sharded, one all-reduce of counters)',
                       'l2': '256 MiB write between timed steps (outside the per-step events); '
                             'the fused kernel reads only ~50 KB of graph tables from HBM',
                       'failures': fails, 'failure_rate': fails / shots, 'wall_s_timed_region': t_wall},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': 88576, 'kernel': 'k_mc_uf',
                         'traffic_note': 'dram__bytes_read.sum + dram__bytes_write.sum of one k_mc_uf launch '
                                         '(ncu --set full, profiles/r01_v5_k_mc_uf_summary.txt): graph tables only',
                         'algorithmic_bytes_per_shot': bps, 'peak_source': peak_src,
                         'note': 'fused kernel keeps a shot in shared memory: the HBM fraction is low by '
                                 'construction; the kernel is bound by shared-memory latency (see profiles/)'},
            'e2e': {'value': e2e_value, 'unit': 'shots/s',
                    'h2d_bytes_per_step': 8 * flat.n_classes + 40, 'd2h_bytes_per_step': 16,
                    'api': 'code.SCHEME.run(decoder, p, n) -> luf_mc_run_host', 'failures': m_e2e},
            'gpu_launches': launches,
            'clocks': clocks.summary(),
        }
        if world == 1 and not args.no_cpu:
            orc = cpu_oracle(code)
            threads = len(os.sched_getaffinity(0))
            rate, n, m, dt = time_cpu(orc, code.NOISE.class_probabilities(p), args.cpu_seconds, threads)
            line['cpu_baseline'] = {'value': rate, 'unit': 'shots/s', 'cores': threads, 'kind': 'port',
                                    'sample': f'{n} shots of the same workload in {dt:.1f} s '
                                              f'(C oracle of the reference algorithm, {m} failures)'}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--shots-per-step', type=int, default=1 << 22)
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu', action='store_true')
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == 'reference':
        run_reference(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
