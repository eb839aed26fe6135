#!/usr/bin/env python
"""Where the time of the host-buffer decode path (luf_decode_batch_host) goes on this box: page-locked copy
bandwidth in both directions, the device-buffer decode of the same shots, the host-buffer call."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import localuf_b200 as L
from localuf_b200 import _engine

code = L.Surface(11, 'phenomenological')
eng = _engine.Engine(code, device=0)
dev = torch.device('cuda', 0)
S = 1 << 20
err = torch.empty((S, eng.WE), dtype=torch.int32, device=dev)
syn = torch.empty((S, eng.WD), dtype=torch.int32, device=dev)
corr = torch.empty((S, eng.WE), dtype=torch.int32, device=dev)
eng.sample(0.01, 1, 0, S, err, syn)
torch.cuda.synchronize()
syn_h = torch.empty(S, eng.WD, dtype=torch.int32).pin_memory()
corr_h = torch.empty(S, eng.WE, dtype=torch.int32).pin_memory()
for name, f, nbytes in (('H2D', lambda: syn.copy_(syn_h, non_blocking=True), syn_h.numel() * 4),
                        ('D2H', lambda: corr_h.copy_(corr, non_blocking=True), corr_h.numel() * 4)):
    f(); torch.cuda.synchronize()
    t0 = time.perf_counter(); f(); torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'{name}: {nbytes / dt / 1e9:.1f} GB/s ({nbytes / 1e6:.0f} MB in {dt * 1e3:.1f} ms)')
eng.sample(0.01, 1, 0, S, err, syn)
syn_h.copy_(syn); torch.cuda.synchronize()
t0 = time.perf_counter(); eng.decode_batch(0, 0, S, syn, corr); torch.cuda.synchronize()
print(f'device-buffer decode: {(time.perf_counter() - t0) * 1e3:.1f} ms')
syn_np, corr_np = syn_h.numpy().view(np.uint32), corr_h.numpy().view(np.uint32)
for n in (1 << 14, S, S, S):
    t0 = time.perf_counter()
    eng.decode_batch_host(0, 0, syn_np[:n], want_steps=False, corr_out=corr_np[:n])
    dt = time.perf_counter() - t0
    print(f'host-buffer decode of {n} shots: {dt * 1e3:.1f} ms = {n / dt / 1e6:.1f} M shots/s')
for chunk in (1 << 16, 1 << 18):
    t0 = time.perf_counter()
    for lo in range(0, S, chunk):
        eng.decode_batch(0, 0, chunk, syn[lo:lo + chunk], corr[lo:lo + chunk])
    torch.cuda.synchronize()
    print(f'device-buffer decode in chunks of {chunk}: {(time.perf_counter() - t0) * 1e3:.1f} ms')
