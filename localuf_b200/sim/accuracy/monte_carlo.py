"""Threshold data: logical failures per (distance, noise level).

Same signature and DataFrame shape as ``localuf/sim/accuracy/monte_carlo.py:25-114``
(rows ``'m'``, ``'n'``; column MultiIndex ``('d', 'p')``); the shots of every
point run inside ``code.SCHEME.run`` on the device.
"""
from __future__ import annotations

from pandas import DataFrame

from localuf_b200.sim._height_calculator import get_heights


def monte_carlo(sample_counts, code_class, decoder_class, noise, scheme='batch',
                get_commit_height=None, get_buffer_height=None, parametrization='balanced',
                demolition=False, monolingual=False, _merge_redundant_edges=True,
                **kwargs_for_decoder_class):
    results = {}
    for d, points in sample_counts.items():
        window_height, commit_height, buffer_height = get_heights(
            d, points[0][1], scheme, get_commit_height, get_buffer_height)
        code = code_class(
            d, noise, scheme=scheme, window_height=window_height, commit_height=commit_height,
            buffer_height=buffer_height, parametrization=parametrization, demolition=demolition,
            monolingual=monolingual, _merge_redundant_edges=_merge_redundant_edges)
        decoder = decoder_class(code, **kwargs_for_decoder_class)
        for noise_level, n in points:
            results[d, noise_level] = code.SCHEME.run(decoder, noise_level, n)
    df = DataFrame(results, index=('m', 'n'))
    df.columns.set_names(['d', 'p'], inplace=True)
    return df


def monte_carlo_results_to_sample_counts(results: DataFrame):
    """Inverse of :func:`monte_carlo`'s first argument (``monte_carlo.py:104-114``)."""
    out: dict = {}
    for (d, noise_level), (_, n) in results.items():
        out.setdefault(d, []).append((noise_level, int(n)))
    return out


def monte_carlo_special(ds, noise_levels, ns, code_class, decoder_class, parametrization='balanced',
                        demolition=False, monolingual=False, _merge_redundant_edges=True,
                        **kwargs_for_decoder_class):
    """Threshold data where the code's noise model is circuit-level and the decoder's is phenomenological
    (``localuf/sim/accuracy/monte_carlo.py:162-201``): same input and output as the reference.  The
    circuit-level graph samples errors and syndromes on the device, the decoder's phenomenological graph
    (same nodes) decodes them, and the leftover parity is taken across the two edge sets."""
    from pandas import MultiIndex
    from localuf_b200 import dist
    if isinstance(ns, int):
        ns = [ns] * len(noise_levels)
    elif len(ns) != len(noise_levels):
        raise ValueError('Sequence ns must have same length as noise_levels')
    mi = MultiIndex.from_product([list(ds), list(noise_levels)], names=['d', 'p'])
    df = DataFrame(columns=mi)
    for d in ds:
        code = code_class(d, noise='circuit-level', parametrization=parametrization, demolition=demolition,
                          monolingual=monolingual, _merge_redundant_edges=_merge_redundant_edges)
        decoder = decoder_class(code_class(d, noise='phenomenological'), **kwargs_for_decoder_class)
        from localuf_b200 import _engine
        noise_engine = _engine.Engine(code)
        dc = {}
        for noise_level, n in zip(noise_levels, ns):
            rank, size = dist.world()
            seed = decoder._draw_seed(None)
            if size > 1:
                seed = dist.allreduce_counts([seed if rank == 0 else 0], device=f'cuda:{decoder.engine.device}')[0]
            lo, cnt = dist.shard(n, rank, size)
            m = decoder.engine.mc_special(noise_engine, decoder._DECODER_ID, decoder._flags, noise_level, seed, lo, cnt)[0] if cnt else 0
            m = dist.allreduce_counts([m], device=f'cuda:{decoder.engine.device}')[0]
            dc[noise_level] = (m, code.SCHEME._slenderness(n))
        df[d] = DataFrame(dc, index=('m', 'n'))
    return df


def subset_sample(ds, noise_level: float, n: int, code_class, decoder_class, noise, parametrization='balanced',
                  tol: float = 5e-1):
    """Threshold data by subset sampling (``localuf/sim/accuracy/monte_carlo.py:204-233``): a
    DataFrame indexed by (distance, error weight) with columns 'subset prob', 'survival prob',
    'm', 'n'; every subset's ``n`` fixed-weight shots run on the device."""
    import pandas as pd
    frames = []
    for d in ds:
        code = code_class(d, noise=noise, parametrization=parametrization)
        frames.append(decoder_class(code).subset_sample(noise_level, n, tol=tol))
    return pd.concat(frames, keys=ds, names=['d'])
