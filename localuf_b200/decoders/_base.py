"""Common host-side behaviour of the device-backed decoders.

Mirrors ``localuf/_base_classes.py:641-712`` (``Decoder``) and the UF-family
state of ``localuf/decoders/_base_uf.py:33-99`` (``growth``, ``erasure``,
``syndrome``): the attributes exist with the reference's names and types, but
they are views of bit arrays that a CUDA kernel produced.
"""
from __future__ import annotations

import random

import numpy as np

from localuf_b200 import _engine, dist
from localuf_b200.constants import Growth, GROWTH_FROM_DEVICE


class Decoder:
    _DECODER_ID: int

    def __init__(self, code):
        self._CODE = code
        self.correction: set = set()
        self._engine = None
        self._flags = 0

    @property
    def CODE(self):
        return self._CODE

    @property
    def engine(self) -> '_engine.Engine':
        """The uploaded graph; created on first use.  Raises
        :class:`localuf_b200._engine.EngineUnavailable` without a GPU -- there is no CPU path."""
        if self._engine is None:
            self._engine = _engine.Engine(self._CODE)
        return self._engine

    def reset(self):
        self.correction = set()

    # seeds: the reference draws from Python's global ``random`` stream, so
    # ``random.seed(s)`` must make a run reproducible here as well.
    @staticmethod
    def _draw_seed(seed):
        return random.getrandbits(63) if seed is None else int(seed)

    def _sharded_counts(self, run_local, n: int, seed):
        """Split ``n`` shots over the ranks of the default process group, run
        the local share, all-reduce ``[failures, shots]``."""
        rank, size = dist.world()
        seed = self._draw_seed(seed)
        if size > 1:
            seed = dist.allreduce_counts([seed if rank == 0 else 0])[0]   # rank 0's seed for everybody
        lo, cnt = dist.shard(n, rank, size)
        fails, shots = run_local(seed, lo, cnt) if cnt else (0, 0)
        fails, shots = dist.allreduce_counts([fails, shots], device=f'cuda:{self.engine.device}')
        assert shots == n
        return fails


    def _mc_forward(self, noise_level: float, n: int, shots: int = 1, seed=None, draw=False, log_history=False,
                    **kwargs):
        """``shots`` independent ``Forward.run(n)`` experiments on the device
        (``_schemes.py:392-450``); returns (logical errors, step counts [shots*n][2])."""
        if draw or log_history:
            raise NotImplementedError("Drawing / history is outside this engine (SURVEY.md section 2, row 14).")
        eng = self.engine
        rank, size = dist.world()
        seed = self._draw_seed(seed)
        if size > 1:
            seed = dist.allreduce_counts([seed if rank == 0 else 0])[0]
        lo, cnt = dist.shard(shots, rank, size)
        counts, steps = ([0, 0, 0, 0], None)
        if cnt:
            counts, steps = eng.forward_mc_host(self._DECODER_ID, self._flags, noise_level, seed, lo, cnt, n,
                                                want_steps=True)
        m = dist.allreduce_counts([counts[0]], device=f'cuda:{eng.device}')[0]
        return m, ([] if steps is None else steps.reshape(-1, 2))


class BaseUF(Decoder):
    """Adds ``growth`` / ``erasure`` / ``syndrome`` (``_base_uf.py:33-99``)."""

    def __init__(self, code):
        super().__init__(code)
        self.syndrome: set = set()
        self._growth_codes = np.zeros(code.N_EDGES, np.uint8)

    @property
    def growth(self) -> dict:
        edges = self._CODE.EDGES
        return {edges[k]: GROWTH_FROM_DEVICE[c] for k, c in enumerate(self._growth_codes.tolist())}

    @property
    def erasure(self) -> set:
        edges = self._CODE.EDGES
        return {edges[k] for k in np.flatnonzero(self._growth_codes == 2).tolist()}

    def reset(self):
        super().reset()
        self.syndrome = set()
        self._growth_codes = np.zeros(self._CODE.N_EDGES, np.uint8)

    def _decode_one(self, syndrome):
        """Decode one syndrome on the device; returns (correction set, steps row)."""
        flat = self._CODE.FLAT
        syn = flat.syndrome_to_bits(syndrome)[None, :]
        corr, growth2b, steps = self.engine.decode_batch_host(
            self._DECODER_ID, self._flags, syn, want_growth=True, want_steps=True)
        self._growth_codes = _engine.unpack_growth(growth2b, flat.n_edges)[0]
        return flat.bits_to_edges(corr[0]), steps[0]


__all__ = ['Decoder', 'BaseUF', 'Growth']
