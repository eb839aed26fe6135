"""Latency data of Snowflake (reference: ``localuf/sim/latency.py:13-72``)."""
from __future__ import annotations

from pandas import DataFrame

from localuf_b200.decoders import Snowflake
from localuf_b200.sim._height_calculator import get_heights


def frugal(ds, noise_levels, n, code_class, noise, time_only='merging', get_commit_height=None,
           get_buffer_height=None, **kwargs_for_Snowflake):
    """Columns ``(d, p)``, one row per latency sample (``Frugal.sample_latency``); the ``n``
    samples of a column come from one kernel launch."""
    data = {}
    for d in ds:
        _, commit_height, buffer_height = get_heights(d, n, 'frugal', get_commit_height, get_buffer_height)
        code = code_class(d, noise, scheme='frugal', window_height=None, commit_height=commit_height,
                          buffer_height=buffer_height)
        decoder = Snowflake(code, **kwargs_for_Snowflake)
        for noise_level in noise_levels:
            data[d, noise_level] = code.SCHEME.sample_latency(decoder, noise_level, time_only=time_only, shots=n)
    df = DataFrame(data).sort_index(axis=1)
    df.columns.set_names(['d', 'p'], inplace=True)
    return df
