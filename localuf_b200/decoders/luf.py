"""Macar and Actis: local Union-Find as a cellular automaton, device-backed.

Same constructors and return values as the reference
(``localuf/decoders/luf/__init__.py:13-41``, ``luf/main.py:83-166``):
``decode`` returns ``(tSV, tBP)`` -- the number of timesteps to validate the
syndrome and to burn-and-peel -- and leaves the correction in ``correction``
(always empty for Actis, whose nodes have no burn/peel action,
``luf/main.py:1014-1023``).  The automaton runs as synchronous warp sweeps in
``csrc/luf_ca.cuh``.
"""
from __future__ import annotations

from localuf_b200 import _engine
from localuf_b200.decoders._base import BaseUF
from localuf_b200.schemes import Batch, Frugal


class _LUF(BaseUF):
    def __init__(self, code):
        if code._CODE_DIMENSION != 2:
            raise TypeError('Local UF decoders are compatible only with the surface code.')
        super().__init__(code)
        self._pending = (set(), 0)

    def decode(self, syndrome, draw=False, log_history=False, **kwargs):
        if draw or log_history:
            raise NotImplementedError("Drawing / history is outside this engine (SURVEY.md section 2, row 14).")
        tSV = self.validate(syndrome)
        tBP = self.peel(False)
        return tSV, tBP

    def validate(self, syndrome, draw=False, log_history=False, **kwargs):
        """``luf/main.py:112-145``: returns tSV; ``growth``/``erasure`` show the validated erasure."""
        if draw or log_history:
            raise NotImplementedError("Drawing / history is outside this engine (SURVEY.md section 2, row 14).")
        self.syndrome = set(syndrome)
        corr, steps = self._decode_one(self.syndrome)
        if steps[0] < 0:
            raise RuntimeError('cellular automaton exceeded the timestep cap')
        self._pending = (corr, int(steps[1]))
        return int(steps[0])

    def peel(self, log_history=False):
        """``luf/main.py:147-166``: returns tBP and publishes the correction."""
        corr, tBP = self._pending
        self.correction = set(corr)
        return tBP

    def reset(self):
        super().reset()
        self._pending = (set(), 0)

    def _mc_batch(self, noise_level: float, n: int, seed=None) -> int:
        eng = self.engine
        return self._sharded_counts(
            lambda s, lo, cnt: eng.mc_run_host(self._DECODER_ID, self._flags, noise_level, s, lo, cnt), n, seed)

    def sample_steps(self, noise_level: float, n: int, seed=None):
        """(tSV, tBP) of ``n`` sampled shots as an int32 array [n, 2] (device: K1 then K3a)."""
        eng = self.engine
        _, syn = eng.sample_host(noise_level, self._draw_seed(seed), 0, n)
        _, _, steps = eng.decode_batch_host(self._DECODER_ID, self._flags, syn)
        return steps


class Macar(_LUF):
    """Controller sees every node (``luf/__init__.py:13-23``)."""
    _DECODER_ID = _engine.DEC_MACAR

    def __init__(self, code):
        if isinstance(code.SCHEME, Frugal):
            raise ValueError('Macar incompatible with frugal scheme.')
        super().__init__(code)


class Actis(_LUF):
    """Strictly local: the controller only talks to node 0 (``luf/__init__.py:26-41``)."""
    _DECODER_ID = _engine.DEC_ACTIS

    def __init__(self, code, _optimal: bool = True):
        if not isinstance(code.SCHEME, Batch):
            raise ValueError('Actis compatible only with batch scheme.')
        super().__init__(code)
        self._flags = 0 if _optimal else _engine.ACTIS_UNOPTIMAL
