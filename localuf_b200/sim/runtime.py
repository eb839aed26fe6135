"""Runtime (timestep-count) data of the cellular-automaton decoders.

Same signatures and DataFrame shapes as ``localuf/sim/runtime.py:25-75``
(``batch``): columns ``(stage, d, p)`` with stage ``'SV'`` / ``'BP'``, one row
per sample.  Samples are drawn and decoded on the device.
"""
from __future__ import annotations

from pandas import DataFrame

from localuf_b200.codes import Surface
from localuf_b200.decoders import Macar


def batch(ds, noise_levels, n, noise, decoder_class=Macar, validate_only=True, **kwargs_for_Surface):
    data = {}
    for d in ds:
        code = Surface(d, noise, scheme='batch', window_height=None, **kwargs_for_Surface)
        decoder = decoder_class(code)
        for noise_level in noise_levels:
            steps = decoder.sample_steps(noise_level, n)
            data['SV', d, noise_level] = steps[:, 0].tolist()
            if not validate_only:
                data['BP', d, noise_level] = steps[:, 1].tolist()
    df = DataFrame(data).sort_index(axis=1)
    df.columns.set_names(['stage', 'd', 'p'], inplace=True)
    return df


def forward(ds, noise_levels, n, noise, get_commit_height=None, get_buffer_height=None, **kwargs_for_Surface):
    """Runtime data of Macar under the forward scheme (``localuf/sim/runtime.py:78-136``)."""
    from localuf_b200.sim._height_calculator import get_heights
    data = {}
    for d in ds:
        _, commit_height, buffer_height = get_heights(d, n, 'forward', get_commit_height, get_buffer_height)
        code = Surface(d, noise, scheme='forward', window_height=None, commit_height=commit_height,
                       buffer_height=buffer_height, **kwargs_for_Surface)
        decoder = Macar(code)
        for noise_level in noise_levels:
            code.SCHEME.run(decoder, noise_level, n)
            data['SV', d, noise_level] = [sv for sv, _ in code.SCHEME.step_counts]
            data['BP', d, noise_level] = [bp for _, bp in code.SCHEME.step_counts]
    df = DataFrame(data).sort_index(axis=1)
    df.columns.set_names(['stage', 'd', 'p'], inplace=True)
    return df


def frugal(ds, noise_levels, n, code_class, noise, time_only='merging', get_commit_height=None,
           get_buffer_height=None, **kwargs_for_Snowflake):
    """Runtime data of Snowflake under the frugal scheme (``localuf/sim/runtime.py:139-199``):
    columns ``(d, p)``, one row per timed block of ``d`` decoding cycles of ``Frugal.run(n)``."""
    from localuf_b200.decoders import Snowflake
    from localuf_b200.sim._height_calculator import get_heights
    data = {}
    for d in ds:
        _, commit_height, buffer_height = get_heights(d, n, 'frugal', get_commit_height, get_buffer_height)
        code = code_class(d, noise, scheme='frugal', window_height=None, commit_height=commit_height,
                          buffer_height=buffer_height)
        decoder = Snowflake(code, **kwargs_for_Snowflake)
        for noise_level in noise_levels:
            code.SCHEME.run(decoder, noise_level, n, time_only=time_only)
            data[d, noise_level] = tuple(code.SCHEME.step_counts)
    df = DataFrame(data).sort_index(axis=1)
    df.columns.set_names(['d', 'p'], inplace=True)
    return df
