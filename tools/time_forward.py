import sys, time
sys.path.insert(0, '/root/repo')
import localuf_b200 as L
for d, noise, p in ((9, 'circuit-level', 0.003), (13, 'circuit-level', 0.003), (13, 'circuit-level', 0.006), (11, 'phenomenological', 0.01), (17, 'circuit-level', 0.004)):
    for name in ('UF', 'Macar'):
        code = L.Surface(d, noise, scheme='forward')
        dec = getattr(L.decoders, name)(code)
        h = code.SCHEME.WINDOW_HEIGHT
        try:
            code.SCHEME.run(dec, p, 1, shots=296, seed=1)
            n = 296
            t0 = time.perf_counter(); code.SCHEME.run(dec, p, 1, shots=n, seed=2); dt = time.perf_counter() - t0
            n = max(n, min(1 << 18, int(n * 0.5 / max(dt, 1e-4))))
            t0 = time.perf_counter(); m, sl = code.SCHEME.run(dec, p, 1, shots=n, seed=3); dt = time.perf_counter() - t0
            print(f'd={d} {noise} p={p} {name}: h={h} N={code.FLAT.n_nodes} E={code.FLAT.n_edges} runs={n} m={m} {n / dt:.0f} runs/s  steps/run={len(code.SCHEME.step_counts) / n:.1f}', flush=True)
        except Exception as e:
            print(d, noise, name, 'ERR', repr(e)[:200], flush=True)
