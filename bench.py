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
    # name: (code class, d, noise, p, description[, scheme, decoder])
    'cfg1': ('Repetition', 5, 'code capacity', 0.05, 'Repetition d=5 code-capacity p=0.05 UF batch'),
    'cfg2': ('Surface', 9, 'code capacity', 0.05, 'Surface d=9 code-capacity p=0.05 UF batch'),
    'cfg3': ('Surface', 11, 'phenomenological', 0.01, 'Surface d=11 phenomenological p=0.01 h=11 UF batch'),
    'cfg3-staged': ('Surface', 11, 'phenomenological', 0.01,
                    'Surface d=11 phenomenological p=0.01 h=11 UF batch, staged pipeline: sample -> HBM -> decode -> HBM -> check'),
    'cfg3-macar': ('Surface', 11, 'phenomenological', 0.01, 'Surface d=11 phenomenological p=0.01 h=11 Macar batch',
                   'batch', 'Macar'),
    'cfg3-actis': ('Surface', 11, 'phenomenological', 0.01, 'Surface d=11 phenomenological p=0.01 h=11 Actis batch',
                   'batch', 'Actis'),
    # BASELINE configs[4] (threshold sweep), two of its hard corners: near threshold at a large distance (shots with odd
    # clusters on both sides go through the merge-log pass and k_hard_resolve), and a graph beyond the 16-bit layout
    # (compact fast launch + the wide instance for its HARD shots)
    'cfg5-d17': ('Surface', 17, 'phenomenological', 0.026, 'Surface d=17 phenomenological p=0.026 h=17 UF batch'),
    'cfg5-wide': ('Surface', 23, 'circuit-level', 0.004, 'Surface d=23 circuit-level p=0.004 h=23 UF batch (65869 edges: wide instance)'),
    'cfg4': ('Surface', 13, 'circuit-level', 0.003,
             'Surface d=13 circuit-level p=0.003 frugal (c=1, b=12) Snowflake; unit = decoding cycle', 'frugal', 'Snowflake'),
}
SEED = 20260101
DECODER_ID = {'UF': 0, 'Macar': 1, 'Actis': 2}


OVERRIDE = {'p': None, 'd': None}      # --p / --d: the same workload at another noise level / distance (sweeps, profiles)


def make_code(name, p_over=None):
    import localuf_b200 as L
    cls, d, noise, p, desc = WORKLOADS[name][:5]
    scheme = WORKLOADS[name][5] if len(WORKLOADS[name]) > 5 else 'batch'
    if p_over is not None:
        desc = desc.replace(f'p={p}', f'p={p_over}')
        p = p_over
    elif OVERRIDE['p'] is not None:
        desc = desc.replace(f'p={p}', f"p={OVERRIDE['p']}")
        p = OVERRIDE['p']
    if OVERRIDE['d'] is not None:
        desc = desc.replace(f'd={d} ', f"d={OVERRIDE['d']} ").replace(f'h={d} ', f"h={OVERRIDE['d']} ")
        d = OVERRIDE['d']
    return getattr(L, cls)(d, noise, scheme=scheme), p, desc


def decoder_name(name):
    return WORKLOADS[name][6] if len(WORKLOADS[name]) > 6 else 'UF'


def algorithmic_bytes_per_shot(flat) -> int:
    """B = 4 W(E) + 2 W(D) bytes (error written and re-read,
    correction written and read, defects written and read), W(x) = 4 ceil(x/32)."""
    we = 4 * ((flat.n_edges + 31) // 32)
    wd = 4 * ((flat.n_detectors + 31) // 32)
    return 4 * we + 2 * wd


def algorithmic_bytes_per_cycle(tables) -> int:
    """per decoding cycle 4 W(fresh edges) + 2 W(new detectors) bytes."""
    return 4 * 4 * ((tables.EL + 31) // 32) + 2 * 4 * ((tables.NLd + 31) // 32)


def measured_peak_hbm():
    path = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    try:
        with open(path) as f:
            return float(json.load(f)['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    except Exception:
        return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler:
    """SM clock and throttle reasons sampled every ~10 ms through NVML (in
    process; `nvidia-smi -lms` as the fallback).  `mark()` is called at the start
    and at the end of the timed region; only samples between the marks count."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')
    NAMES = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')

    def __init__(self, index: int):
        self.rows, self.marks, self.index = [], [], index      # rows: (t, sm_mhz, max_mhz, [reasons])
        self.stop, self.proc, self.thread, self.source = threading.Event(), None, None, None

    def mark(self):
        self.marks.append(time.perf_counter())

    def _physical_index(self):
        vis = os.environ.get('CUDA_VISIBLE_DEVICES')
        if vis:
            ids = [x.strip() for x in vis.split(',') if x.strip()]
            if self.index < len(ids) and ids[self.index].isdigit():
                return int(ids[self.index])
        return self.index

    def _nvml_loop(self, nv, h):
        bits = (('hw_slowdown', nv.nvmlClocksEventReasonHwSlowdown if hasattr(nv, 'nvmlClocksEventReasonHwSlowdown')
                 else nv.nvmlClocksThrottleReasonHwSlowdown),
                ('hw_thermal_slowdown', getattr(nv, 'nvmlClocksEventReasonHwThermalSlowdown',
                                                getattr(nv, 'nvmlClocksThrottleReasonHwThermalSlowdown', 0))),
                ('sw_thermal_slowdown', getattr(nv, 'nvmlClocksEventReasonSwThermalSlowdown',
                                                getattr(nv, 'nvmlClocksThrottleReasonSwThermalSlowdown', 0))),
                ('sw_power_cap', getattr(nv, 'nvmlClocksEventReasonSwPowerCap',
                                         getattr(nv, 'nvmlClocksThrottleReasonSwPowerCap', 0))))
        get_reasons = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mx = float(nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM))
        while not self.stop.is_set():
            try:
                sm = float(nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM))
                r = int(get_reasons(h))
                self.rows.append((time.perf_counter(), sm, mx, [n for n, b in bits if b and (r & b)]))
            except Exception:
                pass
            self.stop.wait(0.01)

    def _smi_loop(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.strip().split(',')]
            if len(parts) < 6:
                continue
            try:
                sm, mx = float(parts[0]), float(parts[1])
            except ValueError:
                continue
            self.rows.append((time.perf_counter(), sm, mx,
                              [n for n, v in zip(self.NAMES, parts[2:6]) if v.lower().startswith('active')]))

    def __enter__(self):
        try:
            import pynvml as nv
            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self._physical_index())
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._nvml_loop, args=(nv, h), daemon=True)
            self.source = 'nvml'
        except Exception:
            try:
                self.proc = subprocess.Popen(
                    ['nvidia-smi', f'--id={self._physical_index()}', f'--query-gpu={self.Q}',
                     '--format=csv,noheader,nounits', '-lms', '20'],
                    stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
                self.thread = threading.Thread(target=self._smi_loop, daemon=True)
                self.source = 'nvidia-smi'
            except Exception:
                self.thread = None
        if self.thread:
            self.thread.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()
        if self.thread:
            self.thread.join(timeout=2)

    def summary(self):
        lo, hi = (self.marks[0], self.marks[-1]) if len(self.marks) >= 2 else (0.0, float('inf'))
        inside = [r for r in self.rows if lo <= r[0] <= hi]
        rows = inside or self.rows          # a timed region shorter than the sampling period: nearest samples
        sm = [r[1] for r in rows]
        reasons = sorted({n for r in rows for n in r[3]})
        return {'sm_mhz': statistics.median(sm) if sm else None, 'sm_max_mhz': max(r[2] for r in rows) if rows else None,
                'reasons': reasons, 'samples': len(inside), 'source': self.source}


def cpu_oracle(code):
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import binding
    return binding.Oracle.from_code(code)


def time_cpu(orc, class_p, target_s, threads, seed=SEED, decoder=0):
    """shots/s of the oracle port on `threads` host threads over ~target_s seconds."""
    n = (2000 if decoder == 0 else 40 if decoder == 1 else 4) * threads
    t0 = time.perf_counter(); orc.mc_batch(decoder, 0, class_p, seed, n, threads); dt = time.perf_counter() - t0
    n = max(n, int(n * target_s / max(dt, 1e-3)))
    t0 = time.perf_counter(); m = orc.mc_batch(decoder, 0, class_p, seed + 1, n, threads); dt = time.perf_counter() - t0
    return n / dt, n, m, dt


def time_cpu_stream(code, class_p, target_s, threads, seed=SEED):
    """decoding cycles/s of the Snowflake/frugal oracle port (Frugal.run(n=1) streams)."""
    sys.path.insert(0, os.path.join(ROOT, 'oracle'))
    import binding
    from localuf_b200 import _window
    orc = binding.StreamOracle(_window.build(code))
    n = 2 * threads
    t0 = time.perf_counter(); orc.mc(class_p, seed, n, 1, threads); dt = time.perf_counter() - t0
    n = max(n, int(n * target_s / max(dt, 1e-3)))
    t0 = time.perf_counter(); m, cycles = orc.mc(class_p, seed + 1, n, 1, threads); dt = time.perf_counter() - t0
    return cycles / dt, n, m, dt


def workload_config(name, desc, code, p):
    """The `config` object, identical in both arms (the driver compares them): what is measured, not how."""
    w = WORKLOADS[name]
    return {'workload': desc, 'code': w[0], 'd': int(code.D), 'noise': w[2], 'p': p,
            'scheme': w[5] if len(w) > 5 else 'batch', 'decoder': decoder_name(name), 'seed': SEED,
            'l2': 'GPU arm: 256 MiB write between timed steps (outside the per-step events), the fused kernels read '
                  'only the graph tables from HBM; CPU arm: not applicable'}


def python_reference_leg(name, p, d, procs, n=None):
    """The UNMODIFIED Python reference on `procs` host cores (baseline/ref_python.py -> baseline/_ref):
    code.SCHEME.run(decoder, p, n) in one process per core (_schemes.py:116-121).  None when the vendored
    reference is absent."""
    sys.path.insert(0, os.path.join(ROOT, 'baseline'))
    try:
        import ref_python
        if not ref_python.available() or name not in ref_python.WORKLOADS:
            return None
        r = ref_python.time_parallel(name, procs, n, p, d)
    except Exception as e:                                   # the baseline must never take the bench line down
        return {'unavailable': str(e)[:200]}
    unit = 'decoding cycles/s' if decoder_name(name) == 'Snowflake' else 'shots/s'
    return {'value': r['value'], 'unit': unit, 'cores': procs, 'kind': 'reference', 'per_core': r['per_core'],
            'sample': f"{r['n_per_process']} = n of code.SCHEME.run(decoder, p, n) in each of {procs} processes "
                      f"({r['units']} units, slowest process {r['seconds']:.1f} s, {r['failures']} failures): the unmodified "
                      'Python reference from baseline/_ref'}


def run_reference(args):
    """--impl reference: the reference's algorithm on the host cores.  The
    reference is pure Python and cannot travel to the GPU box, so this is the
    C restatement in oracle/ (pinned to reference goldens), kind = "port"."""
    rank = int(os.environ.get('RANK', 0))
    if rank != 0:
        return
    code, p, desc = make_code(args.workload)
    dec = decoder_name(args.workload)
    threads = len(os.sched_getaffinity(0))
    class_p = code.NOISE.class_probabilities(p)
    per_step_target = max(2.0, min(20.0, 60.0 / max(args.steps, 1)))
    if dec == 'Snowflake':
        unit = 'decoding cycles/s'
        sys.path.insert(0, os.path.join(ROOT, 'oracle'))
        import binding
        from localuf_b200 import _window
        orc = binding.StreamOracle(_window.build(code))
        for _ in range(args.warmup):
            orc.mc(class_p, 1, threads, 1, threads)
        _, n, _, _ = time_cpu_stream(code, class_p, per_step_target, threads)

        def one(k):
            return orc.mc(class_p, SEED + 10 + k, n, 1, threads)[1]
        what = f'{n} Frugal.run(n=1) memory experiments'
    else:
        unit = 'shots/s'
        orc = cpu_oracle(code)
        did = DECODER_ID[dec]
        for _ in range(args.warmup):
            orc.mc_batch(did, 0, class_p, 1, (500 if did == 0 else 8) * threads, threads)
        _, n, _, _ = time_cpu(orc, class_p, per_step_target, threads, decoder=did)

        def one(k):
            orc.mc_batch(did, 0, class_p, SEED + 10 + k, n, threads)
            return n
        what = f'{n} shots'
    times, units = [], 0
    for k in range(args.steps):
        t0 = time.perf_counter(); units += one(k)
        times.append(time.perf_counter() - t0)
    total = sum(times)
    value = units / total
    print(json.dumps({
        'impl': 'reference', 'metric': METRIC[unit], 'value': value, 'unit': unit,
        'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': 1e3 * total / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': DTYPE, 'data': 'synthetic',
        'config': workload_config(args.workload, desc, code, p),
        'run': {'units_per_step': units // max(args.steps, 1),
                'sampler': 'one Bernoulli draw per edge, xoshiro256** per thread (the GPU arm draws the same '
                           'distribution by geometric skipping over Philox4x32-10)'},
        'cpu_baseline': {'value': value, 'unit': unit, 'cores': threads, 'kind': 'port',
                         'sample': f'{what} per step on {threads} threads (C restatement of the reference algorithm, '
                                   'oracle/luf_oracle*.c, pinned to the reference goldens)'},
        'cpu_baseline_python': None if args.no_python_ref else python_reference_leg(args.workload, args.p, args.d, threads),
        'e2e': {'value': value, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }), file=_RESULT_OUT, flush=True)


_RESULT_OUT = sys.stdout          # main() points it at the real stdout and sends fd 1 to stderr
DTYPE = 'u32 bit-packed / u16 indices'
METRIC = {'shots/s': 'decoded shots/s', 'decoding cycles/s': 'decoding cycles/s'}
DEFAULT_UNITS = {'cfg1': 1 << 26, 'cfg2': 1 << 24, 'cfg3': 1 << 22, 'cfg3-staged': 1 << 20, 'cfg3-macar': 1 << 19, 'cfg3-actis': 1 << 14,
                 'cfg4': 1 << 16, 'cfg5-d17': 1 << 17, 'cfg5-wide': 1 << 15}
# warp-instructions per unit and lanes per instruction of the dominant kernel (ncu --set full of the same workload)
ISSUE = {'cfg3': (3071.0, 15.04, 'profiles/r02_v3_k_mc_uf_fast_summary.txt (3.22 G warp-instructions per 2^20 shots)')}
# dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel (ncu --set full; profiles/)
TRAFFIC = {'cfg3': (76288, 'profiles/r02_v3_k_mc_uf_fast_summary.txt (one launch of 2^20 shots: the graph tables once per block; no per-shot HBM traffic)'),
           'cfg3-macar': (89856, 'profiles/r01_k_mc_ca_macar_summary.txt'),
           'cfg4': (145920, 'profiles/r01_k_sf_mc_acp_summary.txt (warp tile; the block tile reads the same tables)')}


def host_buffer_e2e(eng, p, world, rank, steps, n=1 << 20):
    """The staged path a hardware-in-the-loop caller uses: HOST syndromes in, HOST corrections out through
    luf_decode_batch_host (page-locked buffers, host<->device copies inside the timed region, chunks pipelined
    over three streams).  Returns the `e2e`-style object (shots/s over all ranks)."""
    import numpy as np
    import torch
    import torch.distributed as dist
    WE, WD = eng.WE, eng.WD
    err = torch.empty(n, WE, dtype=torch.int32, device='cuda')
    syn = torch.empty(n, WD, dtype=torch.int32, device='cuda')
    eng.sample(p, SEED + 7, rank * n, n, err, syn)
    syn_h = torch.empty(n, WD, dtype=torch.int32).pin_memory()
    corr_h = torch.empty(n, WE, dtype=torch.int32).pin_memory()
    syn_h.copy_(syn)
    torch.cuda.synchronize()
    syn_np, corr_np = syn_h.numpy().view(np.uint32), corr_h.numpy().view(np.uint32)
    eng.decode_batch_host(0, 0, syn_np, want_steps=False, corr_out=corr_np)       # warm-up: device scratch, streams
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        t1 = time.perf_counter()
        eng.decode_batch_host(0, 0, syn_np, want_steps=False, corr_out=corr_np)
        print(f'[bench] host-buffer decode of {n} shots: {1e3 * (time.perf_counter() - t1):.1f} ms', file=sys.stderr)
    if world > 1:
        dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out = {'value': n * world * steps / float(t.item()), 'unit': 'shots/s', 'h2d_bytes_per_step': 4 * WD * n,
           'd2h_bytes_per_step': 4 * WE * n, 'shots_per_step_per_gpu': n,
           'api': 'Engine.decode_batch_host(UF, syndromes) -> luf_decode_batch_host: host syndromes in, host '
                  'corrections out (page-locked buffers, copies inside the timed region)'}
    # the same with logical frames out (luf_decode_frames_host): one byte per shot leaves the device
    frame_h = torch.empty(n, dtype=torch.uint8).pin_memory()
    frame_np = frame_h.numpy()
    eng.decode_frames_host(0, 0, syn_np, frame_out=frame_np)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(steps):
        eng.decode_frames_host(0, 0, syn_np, frame_out=frame_np)
    if world > 1:
        dist.barrier()
    t = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out['frames'] = {'value': n * world * steps / float(t.item()), 'unit': 'shots/s', 'h2d_bytes_per_step': 4 * WD * n,
                     'd2h_bytes_per_step': n,
                     'api': 'Engine.decode_frames_host(UF, syndromes) -> luf_decode_frames_host: host syndromes in, one '
                            'logical-frame byte per shot out'}
    return out


# the other BASELINE / SURVEY 8d measurement points, timed briefly next to the headline (N = 1 only):
# (workload, noise level or None = the workload's own, units per step)
SIDE_WORKLOADS = [('cfg1', None, 1 << 24), ('cfg2', None, 1 << 22), ('cfg2', 0.10, 1 << 21),
                  ('cfg3', 0.005, 1 << 21), ('cfg3', 0.02, 1 << 20), ('cfg3', 0.026, 1 << 19),
                  ('cfg3-macar', None, 1 << 18), ('cfg3-actis', None, 1 << 13), ('cfg4', None, 1 << 14), ('cfg4', 0.001, 1 << 14),
                  ('cfg5-d17', None, 1 << 16), ('cfg5-wide', None, 1 << 14)]


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
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')     # > 126 MB L2
    peak, peak_src = measured_peak_hbm()
    threads = len(os.sched_getaffinity(0))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def measure(name, S, steps, warmup, p_over=None, e2e_steps=3, host_buffers=False, cpu_seconds=0.0):
        """One workload: `steps` timed steps of S units per GPU (CUDA events per step, max over ranks), the
        same through the public API with host arguments (e2e), roofline and CPU port baseline."""
        code, p, desc = make_code(name, p_over)
        dec = decoder_name(name)
        decoder = getattr(L.decoders, dec)(code)
        counts = torch.zeros(4, dtype=torch.int64, device='cuda')
        stream_mode = dec == 'Snowflake'
        if stream_mode:
            # a step = S independent Frugal.run(n=1) memory experiments per GPU; unit = decoding cycle
            eng = _engine.StreamEngine(code, device=local_rank)
            decoder._engine = eng
            unit, kernel = 'decoding cycles/s', 'k_sf_mc_cta / k_sf_mc'
            bpu = algorithmic_bytes_per_cycle(eng.tables)
            h2d = 8 * eng.tables.n_classes + 40

            def step(k):
                eng.mc(decoder._flags, p, SEED, (k * world + rank) * S, S, 1, counts)

            def e2e_run(n_units, seed):
                return code.SCHEME.run(decoder, p, 1, shots=n_units, seed=seed)[0]
            api = 'code.SCHEME.run(decoder, p, n=1, shots=S) -> luf_stream_mc_host'
        else:
            eng = _engine.Engine(code, device=local_rank)
            decoder._engine = eng
            unit = 'shots/s'
            kernel = {'UF': 'k_mc_uf_fast (+ its merge-log pass and k_hard_resolve over the queued shots)', 'Macar': 'k_mc_ca', 'Actis': 'k_mc_ca'}[dec]
            bpu = algorithmic_bytes_per_shot(code.FLAT)
            h2d = 8 * code.FLAT.n_classes + 40
            did = DECODER_ID[dec]

            def step(k):
                # rank r, step k owns global shots [(k*world + r)*S, +S)
                eng.mc_run(did, 0, p, SEED, (k * world + rank) * S, S, counts)

            def e2e_run(n_units, seed):
                return code.SCHEME.run(decoder, p, n_units, seed=seed)[0]
            api = 'code.SCHEME.run(decoder, p, n) -> luf_mc_run_host'

        for k in range(warmup):
            step(k)
        barrier()
        counts.zero_()
        launches0 = eng.launch_count
        ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
        with ClockSampler(local_rank) as clocks:
            barrier()
            clocks.mark()
            t_wall0 = time.perf_counter()
            for k in range(steps):
                flush.fill_(k & 0xFF)                       # L2 flush between timed steps
                ev[k][0].record()
                step(warmup + k)
                ev[k][1].record()
            barrier()
            t_wall = time.perf_counter() - t_wall0
            clocks.mark()
        launches = eng.launch_count - launches0
        step_ms = [a.elapsed_time(b) for a, b in ev]
        total_ms = torch.tensor([sum(step_ms)], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM)
        total_ms = float(total_ms.item())
        fails, units = (int(x) for x in counts.tolist()[:2])
        if not stream_mode:
            assert units == S * steps * world, (units, S, steps, world)
        value = units / (total_ms / 1e3)
        units_per_step_gpu = units // (steps * world)

        # ---- end to end through the public API (the call a reference user makes):
        # host arguments in (class probabilities, seed, counts), counters back on the host
        for _ in range(2):
            e2e_run(max(1, min(S, 1 << 12)) * world, 1)
        barrier()
        t0 = time.perf_counter()
        e2e_steps = max(1, min(steps, e2e_steps))
        m_e2e = 0
        for k in range(e2e_steps):
            m_e2e += e2e_run(S * world, SEED + 1000 + k)
        barrier()
        t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
        if world > 1:
            dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
        e2e_value = units_per_step_gpu * world * e2e_steps / float(t_e2e.item())

        # ---- and with real host buffers (UF workloads): host syndromes -> host corrections, staged decode kernel
        e2e_host = host_buffer_e2e(eng, p, world, rank, e2e_steps) if host_buffers and dec == 'UF' else None

        kern_ms = statistics.mean(step_ms)
        achieved = bpu * units_per_step_gpu / (kern_ms / 1e3) / 1e9
        traffic, traffic_src = TRAFFIC.get(name, (None, None)) if p_over is None else (None, None)
        # the ceiling that binds the fused UF kernel is the SMs' instruction issue rate, not HBM: warp-instructions per
        # shot and lanes per instruction from the committed ncu capture of the same kernel and workload
        issue = None
        if name in ISSUE and p_over is None and OVERRIDE['p'] is None and OVERRIDE['d'] is None:
            wi, lanes, src = ISSUE[name]
            sm_count = torch.cuda.get_device_properties(torch.cuda.current_device()).multi_processor_count
            mhz = (clocks.summary() or {}).get('sm_mhz') or 1965.0
            peak_wi = sm_count * 4 * mhz * 1e6                       # four schedulers per SM, one warp-instruction per cycle each
            ach_wi = wi * units_per_step_gpu / (kern_ms / 1e3)
            issue = {'unit': 'warp-instructions/s', 'achieved': ach_wi, 'peak': peak_wi, 'frac': ach_wi / peak_wi,
                     'warp_instructions_per_unit': wi, 'lanes_per_instruction': lanes,
                     'thread_instruction_frac': ach_wi / peak_wi * lanes / 32.0, 'source': src}
        r = {
            'metric': METRIC[unit], 'value': value, 'unit': unit, 'n_gpus': world,
            'steps': steps, 'warmup': warmup, 'ms_per_step': total_ms / steps,
            'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
            'dtype': DTYPE, 'data': 'synthetic',
            'config': workload_config(name, desc, code, p),
            'run': {'units_per_step_per_gpu': units_per_step_gpu,
                    'parallelism': f'dp{world} (shots sharded, one all-reduce of counters)',
                    'failures': fails, 'failure_rate': fails / max(units, 1), 'wall_s_timed_region': t_wall},
            'roofline': {'bound': 'hbm', 'achieved': achieved, 'peak': peak, 'unit': 'GB/s',
                         'frac': achieved / peak, 'traffic': traffic, 'kernel': kernel,
                         'traffic_source': traffic_src,
                         'algorithmic_bytes_per_unit': bpu, 'peak_source': peak_src,
                         'note': 'the fused kernels keep a shot in shared memory, so the HBM fraction is low by '
                                 'construction (SURVEY 8d): they are bound by shared-memory latency and issue slots '
                                 '(profiles/)',
                         'issue_slots': issue},
            'e2e': {'value': e2e_value, 'unit': unit, 'h2d_bytes_per_step': h2d, 'd2h_bytes_per_step': 32,
                    'api': api, 'failures': m_e2e},
            'e2e_host_buffers': e2e_host,
            'gpu_launches': launches,
            'clocks': clocks.summary(),
        }
        if rank == 0 and world == 1 and cpu_seconds > 0:
            class_p = code.NOISE.class_probabilities(p)
            if stream_mode:
                rate, n, m, dt = time_cpu_stream(code, class_p, cpu_seconds, threads)
                what = f'{n} Frugal.run(n=1) memory experiments'
            else:
                rate, n, m, dt = time_cpu(cpu_oracle(code), class_p, cpu_seconds, threads, decoder=DECODER_ID[dec])
                what = f'{n} shots'
            r['cpu_baseline'] = {'value': rate, 'unit': unit, 'cores': threads, 'kind': 'port',
                                 'sample': f'{what} of the same workload in {dt:.1f} s '
                                           f'(C restatement of the reference algorithm, oracle/, {m} failures)'}
        del eng, decoder
        return r

    S = args.shots_per_step or DEFAULT_UNITS[args.workload]
    line = measure(args.workload, S, args.steps, args.warmup, host_buffers=True,
                   cpu_seconds=0.0 if args.no_cpu else args.cpu_seconds)
    if rank == 0 and world == 1 and not args.no_python_ref:
        line['cpu_baseline_python'] = python_reference_leg(args.workload, args.p, args.d, threads)
    if world == 1 and not args.no_workloads and args.workload == 'cfg3' and args.p is None and args.d is None:
        side = []
        for name, p_over, units in SIDE_WORKLOADS:
            try:
                r = measure(name, units, 2, 3, p_over=p_over, e2e_steps=1, cpu_seconds=0.0 if args.no_cpu else 1.5)
            except Exception as e:                       # a side measurement must not take the headline down
                side.append({'workload': name, 'p': p_over, 'error': str(e)[:200]})
                continue
            side.append({k: r[k] for k in ('metric', 'value', 'unit', 'ms_per_step', 'config', 'run', 'roofline', 'e2e',
                                          'gpu_launches', 'clocks', 'cpu_baseline') if k in r})
        line['workloads'] = side
    if rank == 0:
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def run_staged(args):
    """The staged pipeline of SURVEY 8d on device buffers -- K1 luf_sample writes bit-packed errors and
    defects to HBM, K2 luf_decode_batch reads the defects and writes corrections, K4 luf_check reads errors
    and corrections -- with every kernel timed by its own CUDA events, and, as `e2e`, the decode of HOST
    syndromes into HOST corrections through luf_decode_batch_host (page-locked buffers, copies inside the
    timed region, chunks pipelined over three streams)."""
    import numpy as np
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
    eng = _engine.Engine(code, device=local_rank)
    S = args.shots_per_step or DEFAULT_UNITS[args.workload]
    WE, WD = eng.WE, eng.WD
    i32 = dict(dtype=torch.int32, device='cuda')
    err, syn, corr = torch.empty(S, WE, **i32), torch.empty(S, WD, **i32), torch.empty(S, WE, **i32)
    logical = torch.empty(S, dtype=torch.uint8, device='cuda')
    fails = torch.zeros(1, dtype=torch.int64, device='cuda')
    flush = torch.empty(256 << 20, dtype=torch.uint8, device='cuda')
    names = ('k_sample', 'k_decode_uf', 'k_check')
    bytes_per_shot = {'k_sample': 4 * (WE + WD), 'k_decode_uf': 4 * (WD + WE), 'k_check': 8 * WE + 1}

    def step(k, ev=None):
        marks = ev or [None] * 4
        if ev: marks[0].record()
        eng.sample(p, SEED, (k * world + rank) * S, S, err, syn)
        if ev: marks[1].record()
        eng.decode_batch(0, 0, S, syn, corr)
        if ev: marks[2].record()
        eng.check(S, err, corr, logical, fails)
        if ev: marks[3].record()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for k in range(args.warmup):
        step(k)
    barrier()
    fails.zero_()
    launches0 = eng.launch_count
    evs = [[torch.cuda.Event(enable_timing=True) for _ in range(4)] for _ in range(args.steps)]
    with ClockSampler(local_rank) as clocks:
        barrier()
        clocks.mark()
        for k in range(args.steps):
            flush.fill_(k & 0xFF)
            step(args.warmup + k, evs[k])
        barrier()
        clocks.mark()
    launches = eng.launch_count - launches0
    kern_ms = {n: statistics.mean(e[i].elapsed_time(e[i + 1]) for e in evs) for i, n in enumerate(names)}
    total_ms = torch.tensor([sum(e[0].elapsed_time(e[3]) for e in evs)], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
        dist.all_reduce(fails, op=dist.ReduceOp.SUM)
    total_ms = float(total_ms.item())
    value = S * args.steps * world / (total_ms / 1e3)

    # ---- end to end: host syndromes in, host corrections out (page-locked), failures counted on the host side
    n_e2e = S
    syn_h = torch.empty(n_e2e, WD, dtype=torch.int32).pin_memory()
    corr_h = torch.empty(n_e2e, WE, dtype=torch.int32).pin_memory()
    syn_h.copy_(syn[:n_e2e])
    torch.cuda.synchronize()
    syn_np, corr_np = syn_h.numpy().view(np.uint32), corr_h.numpy().view(np.uint32)
    eng.decode_batch_host(0, 0, syn_np[:1 << 14], want_steps=False, corr_out=corr_np[:1 << 14])
    barrier()
    e2e_steps = max(1, min(args.steps, 3))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        eng.decode_batch_host(0, 0, syn_np, want_steps=False, corr_out=corr_np)
    barrier()
    t_e2e = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device='cuda')
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_value = n_e2e * world * e2e_steps / float(t_e2e.item())
    same = bool((corr_h.cuda() == corr[:n_e2e]).all().item())     # the host path returns what the device path wrote

    if rank == 0:
        peak, peak_src = measured_peak_hbm()
        kernels = {}
        for n in names:
            gbs = bytes_per_shot[n] * S / (kern_ms[n] / 1e3) / 1e9
            kernels[n] = {'ms': kern_ms[n], 'algorithmic_bytes_per_unit': bytes_per_shot[n], 'achieved': gbs,
                          'frac': gbs / peak}
        dom = max(names, key=lambda n: kern_ms[n])
        bpu = algorithmic_bytes_per_shot(code.FLAT)
        line = {
            'metric': METRIC['shots/s'], 'value': value, 'unit': 'shots/s', 'n_gpus': world, 'steps': args.steps,
            'warmup': args.warmup, 'ms_per_step': total_ms / args.steps, 'higher_is_better': True, 'scaling': 'weak',
            'vs_baseline': None, 'dtype': DTYPE, 'data': 'synthetic',
            'config': workload_config(args.workload, desc, code, p),
            'run': {'units_per_step_per_gpu': S, 'parallelism': f'dp{world} (shots sharded, no data-path collective)',
                    'l2': '256 MiB write between timed steps; the buffers of one step (1.1 GB) exceed the L2',
                    'failures': int(fails.item()), 'host_path_equals_device_path': same},
            'roofline': {'bound': 'hbm', 'achieved': bpu * S / (kern_ms[dom] / 1e3) / 1e9, 'peak': peak, 'unit': 'GB/s',
                         'frac': bpu * S / (kern_ms[dom] / 1e3) / 1e9 / peak, 'traffic': None, 'kernel': dom,
                         'algorithmic_bytes_per_unit': bpu, 'peak_source': peak_src, 'kernels': kernels,
                         'note': 'achieved = the pipeline figure B = 4 W(E) + 2 W(D) per shot over the dominant '
                                 "kernel's time; `kernels` holds every stage with its own bytes and time"},
            'e2e': {'value': e2e_value, 'unit': 'shots/s', 'h2d_bytes_per_step': 4 * WD * n_e2e,
                    'd2h_bytes_per_step': 4 * WE * n_e2e,
                    'api': 'Engine.decode_batch_host(UF, syndromes) -> luf_decode_batch_host (page-locked host buffers)'},
            'gpu_launches': launches,
            'clocks': clocks.summary(),
        }
        print(json.dumps(line), file=_RESULT_OUT, flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='ours', choices=['ours', 'reference'])
    ap.add_argument('--workload', default='cfg3', choices=sorted(WORKLOADS))
    ap.add_argument('--shots-per-step', type=int, default=0, help='units per step per GPU (0 = workload default)')
    ap.add_argument('--cpu-seconds', type=float, default=12.0)
    ap.add_argument('--no-cpu', action='store_true')
    ap.add_argument('--no-python-ref', action='store_true', help='skip the unmodified Python reference leg')
    ap.add_argument('--no-workloads', action='store_true', help='skip the short runs of the other BASELINE configs')
    ap.add_argument('--p', type=float, default=None, help='noise level override')
    ap.add_argument('--d', type=int, default=None, help='code distance override')
    args = ap.parse_args()
    OVERRIDE['p'], OVERRIDE['d'] = args.p, args.d
    if args.warmup < 3:
        args.warmup = 3
    # stdout carries exactly ONE line, the JSON result: whatever libraries write to file descriptor 1 meanwhile
    # (NCCL's version banner under torchrun, for one) goes to stderr
    global _RESULT_OUT
    sys.stdout.flush()
    _RESULT_OUT = os.fdopen(os.dup(1), 'w')
    os.dup2(2, 1)
    if args.impl == 'reference':
        run_reference(args)
    elif args.workload.endswith('-staged'):
        run_staged(args)
    else:
        run_ours(args)


if __name__ == '__main__':
    main()
