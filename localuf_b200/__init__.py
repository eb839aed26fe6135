"""localuf_b200 -- B200-native batched Union-Find decoding engine.

Drop-in for the Monte-Carlo hot path of timchan0/localuf (sample noise ->
syndrome -> decode -> logical-failure count); see DESIGN.md.
"""
from localuf_b200.codes import Repetition, Surface  # noqa: F401
from localuf_b200 import decoders  # noqa: F401
from localuf_b200 import sim  # noqa: F401

__all__ = ['Repetition', 'Surface', 'decoders', 'sim']
