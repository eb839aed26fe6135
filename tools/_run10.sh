timeout 600 python -m pytest tests/test_gpu_dcs.py -x -q 2>&1 | tail -2
python - <<'PY'
import time, torch
import localuf_b200 as L
for d, p in ((11, 0.01), (17, 0.02)):
    code = L.Surface(d, 'phenomenological')
    dec = L.decoders.UF(code)
    dec.sample_confidence_scores(p, 4096, noise_level_for_priors=p, seed=1)
    n = 1 << 19 if d == 11 else 1 << 15
    torch.cuda.synchronize(); t0 = time.perf_counter()
    out = dec.sample_confidence_scores(p, n, noise_level_for_priors=p, seed=2)
    torch.cuda.synchronize(); dt = time.perf_counter() - t0
    print(f'd={d} p={p}: {n / dt / 1e6:.3f} M scored shots/s; mean swim {out["swim_distance"].mean():.3f}, '
          f'uef {out["unclustered_edge_fraction"].mean():.4f}, fail rate {out["logical"].mean():.2e}')
PY
