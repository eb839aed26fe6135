"""Decoding schemes: how shots are cut out of the space-time history and how a
logical failure is read off the leftover.

Host mirror of ``localuf/_schemes.py`` (``Scheme`` ``:25-94``, ``Batch``
``:97-163``, ``Global`` ``:166-203``, ``_Streaming`` ``:206-279``, ``Forward``
``:282-477``, ``Frugal`` ``:480-698``), ``localuf/_pairs.py`` and
``localuf/_determinants.py``.

The set-based helpers (``get_logical_error``, :class:`Pairs`,
:class:`LogicalCounter`) are kept for the drop-in single-shot API and the
host-logic tests.  ``run`` -- the batched Monte-Carlo entry point that
``sim.accuracy.monte_carlo`` calls (``sim/accuracy/monte_carlo.py:98``) -- does
not loop in Python: it hands ``(p, n)`` to the decoder's device engine, which
samples, decodes and counts failures for all shots inside CUDA kernels.
"""
from __future__ import annotations

import abc
import math

Node = tuple
Edge = tuple


# ---------------------------------------------------------------- free strings
class Pairs:
    """End-point pairs of the open error strings seen so far (``_pairs.py:4-66``).

    Stored as a symmetric map ``endpoint -> other endpoint``.
    """

    def __init__(self) -> None:
        self._dc: dict[Node, Node] = {}

    def __repr__(self) -> str:
        return str(self.as_set)

    def __contains__(self, u: Node) -> bool:
        return u in self._dc

    def __getitem__(self, u: Node) -> Node:
        return self._dc[u]

    @property
    def as_set(self) -> set[Edge]:
        seen: set[Node] = set()
        out: set[Edge] = set()
        for u, v in self._dc.items():
            if u not in seen:
                out.add((u, v))
                seen.add(v)
        return out

    def reset(self) -> None:
        self._dc.clear()

    def add(self, u: Node, v: Node) -> None:
        self._dc[u] = v
        self._dc[v] = u

    def remove(self, u: Node) -> None:
        v = self._dc.pop(u)
        del self._dc[v]

    def load(self, e: Edge) -> None:
        """Splice edge ``e`` into the strings: an endpoint that already ends a
        string extends it; two strings meeting at ``e`` are joined; a string
        closed on itself disappears."""
        u, v = e
        far_u = self._dc.get(u)
        far_v = self._dc.get(v)
        if far_u is not None:
            self.remove(u)
            if v in self._dc:           # (re-read: removing u's pair may have removed v's)
                far_v = self._dc[v]
                self.remove(v)
                self.add(far_u, far_v)
            elif v != far_u:
                self.add(v, far_u)
        elif far_v is not None:
            self.remove(v)
            self.add(u, far_v)
        else:
            self.add(u, v)


class LogicalCounter:
    """Counts strings that join opposite boundaries and lowers the rest
    (``_pairs.py:69-111``)."""

    def __init__(self, d: int, commit_height: int, long_axis: int, time_axis: int) -> None:
        self._D, self._COMMIT_HEIGHT = d, commit_height
        self._LONG_AXIS, self._TIME_AXIS = long_axis, time_axis

    def _lower_node(self, v: Node) -> Node:
        w = list(v)
        w[self._TIME_AXIS] -= self._COMMIT_HEIGHT
        return tuple(w)

    def count(self, pairs: Pairs):
        a, d = self._LONG_AXIS, self._D
        errors = 0
        kept = Pairs()
        for u, v in pairs.as_set:
            gap = abs(u[a] - v[a])
            if gap == d:
                errors += 1
            elif gap == 0 and u[a] in (-1, d - 1):
                pass                      # both ends on the same space boundary: trivial
            else:
                kept.load((self._lower_node(u), self._lower_node(v)))
        return errors, kept


# --------------------------------------------------------------- determinants
class _SpaceBoundary:
    """``_determinants.py:4-36``."""

    def __init__(self, d, long_axis):
        self.D, self.LONG_AXIS = d, long_axis

    def is_boundary(self, v):
        x = v[self.LONG_AXIS]
        return x == -1 or x == self.D - 1


class _SpaceTimeBoundary(_SpaceBoundary):
    """``_determinants.py:39-71``: the top sheet ``t == h`` is a boundary too."""

    def __init__(self, d, long_axis, time_axis, window_height):
        super().__init__(d, long_axis)
        self.TIME_AXIS, self.WINDOW_HEIGHT = time_axis, window_height

    def is_boundary(self, v):
        return super().is_boundary(v) or v[self.TIME_AXIS] == self.WINDOW_HEIGHT


# -------------------------------------------------------------------- schemes
class Scheme(abc.ABC):
    def __init__(self, code):
        self._CODE = code

    @property
    @abc.abstractmethod
    def WINDOW_HEIGHT(self) -> int: ...

    @abc.abstractmethod
    def get_logical_error(self, leftover) -> int: ...

    def is_boundary(self, v: Node) -> bool:
        return self._DETERMINANT.is_boundary(v)

    @abc.abstractmethod
    def run(self, decoder, noise_level: float, n: int, **kwargs): ...

    def sim_cycles_given_weight(self, decoder, weight, n):
        raise NotImplementedError(f"Subset sampling is defined for the batch scheme only, not {self}.")


class Batch(Scheme):
    """One shot = one decoding graph of ``h`` layers (``_schemes.py:97-163``)."""

    def __str__(self) -> str:
        return 'batch'

    def __init__(self, code, window_height: int):
        super().__init__(code)
        self._WINDOW_HEIGHT = window_height
        self._DETERMINANT = _SpaceBoundary(code.D, code.LONG_AXIS)

    @property
    def WINDOW_HEIGHT(self): return self._WINDOW_HEIGHT

    def get_logical_error(self, leftover) -> int:
        """Parity of the leftover edges that start on the west boundary."""
        a = self._CODE.LONG_AXIS
        return sum(1 for u, _ in leftover if u[a] == -1) & 1

    def _slenderness(self, n):
        if str(self._CODE.NOISE) == 'code capacity':
            return n
        return self.WINDOW_HEIGHT * n / self._CODE.D

    def run(self, decoder, noise_level: float, n: int, **kwargs):
        """``n`` shots of sample -> syndrome -> decode -> logical parity, all on
        the device; returns ``(failures, slenderness)`` like ``_schemes.py:116-121``."""
        m = decoder._mc_batch(noise_level, n, **kwargs)
        return m, self._slenderness(n)

    def sim_cycles_given_weight(self, decoder, weight, n: int, **kwargs):
        """``n`` shots whose error has exactly weight ``weight`` (``_schemes.py:157-163``);
        returns ``(failures, n)``.  Uniform noise models: sampled on the device
        (``luf_sample_weight``); circuit-level weights (tuples): ``Noise.force_error`` on the
        host, decoding and the logical check on the device."""
        return decoder._mc_weight(weight, n, **kwargs), n

    def _sim_cycle_given_error(self, decoder, error) -> int:
        """Single shot through the drop-in API (``_schemes.py:143-155``)."""
        syndrome = self._CODE.get_syndrome(error)
        decoder.reset()
        decoder.decode(syndrome)
        return self.get_logical_error(set(error) ^ decoder.correction)

    def _sim_cycle_given_p(self, decoder, noise_level) -> int:
        return self._sim_cycle_given_error(decoder, self._CODE.make_error(noise_level))


class Global(Batch):
    """One tall window of ``d*n`` layers; logical error strings are *counted*
    (``_schemes.py:166-203``)."""

    def __str__(self) -> str:
        return 'global batch'

    def __init__(self, code, window_height: int):
        self.pairs = Pairs()
        super().__init__(code, window_height)

    def reset(self):
        self.pairs.reset()

    def get_logical_error(self, leftover) -> int:
        """``_schemes.py:194-203``.  Where three or more leftover edges meet, the pairing of
        string ends depends on the order of the loads; the reference iterates a ``set``,
        here the order is ascending edge id (``oracle/gen_golden_gl.py`` stores both counts)."""
        index = self._CODE.FLAT.edge_index
        for e in sorted(leftover, key=index.__getitem__):
            self.pairs.load(e)
        a, d = self._CODE.LONG_AXIS, self._CODE.D
        return sum(1 for u, v in self.pairs.as_set if abs(u[a] - v[a]) == d)

    def count_leftover_bits(self, leftover_bits) -> int:
        """The same for one bit-packed leftover (``err ^ corr`` of the device)."""
        import numpy as np
        bits = np.unpackbits(np.ascontiguousarray(leftover_bits, dtype=np.uint32).view(np.uint8), bitorder='little')
        edges = self._CODE.EDGES
        self.reset()
        for k in np.flatnonzero(bits[:len(edges)]).tolist():
            self.pairs.load(edges[k])
        a, d = self._CODE.LONG_AXIS, self._CODE.D
        return sum(1 for u, v in self.pairs.as_set if abs(u[a] - v[a]) == d)

    def run(self, decoder, noise_level: float, n: int, shots: int = 1, **kwargs):
        """``_schemes.py:186-191``: ONE window of ``d*n`` layers is sampled and decoded on the
        device (K1, K2/K3a), its leftover counted here; ``shots`` > 1 repeats that with
        independent noise and sums (the reference always runs one)."""
        assert self.WINDOW_HEIGHT == self._CODE.D * n
        self.reset()
        return decoder._mc_global(noise_level, shots=shots, **kwargs), n * shots


class _Streaming(Scheme):
    def __init__(self, code, commit_height, buffer_height, commit_edges):
        super().__init__(code)
        self._COMMIT_HEIGHT = commit_height
        self._BUFFER_HEIGHT = buffer_height
        self._COMMIT_EDGES = tuple(commit_edges)
        in_commit = set(commit_edges)
        self._BUFFER_EDGES = tuple(e for e in code.EDGES if e not in in_commit)
        self._DETERMINANT = _SpaceTimeBoundary(
            code.D, code.LONG_AXIS, code.TIME_AXIS, commit_height + buffer_height)
        self.pairs = Pairs()
        self._LOGICAL_COUNTER = LogicalCounter(
            code.D, commit_height, code.LONG_AXIS, code.TIME_AXIS)
        self.step_counts: list = []

    @property
    def WINDOW_HEIGHT(self): return self._COMMIT_HEIGHT + self._BUFFER_HEIGHT

    def reset(self):
        self.pairs.reset()
        self.step_counts.clear()

    def _count_and_lower(self) -> int:
        count, self.pairs = self._LOGICAL_COUNTER.count(self.pairs)
        return count


class Forward(_Streaming):
    """Sliding window committing ``c`` layers per decode (``_schemes.py:282-450``)."""

    def __str__(self) -> str:
        return 'forward'

    def get_logical_error(self, commit_leftover) -> int:
        for e in commit_leftover:
            self.pairs.load(e)
        return self._count_and_lower()

    def _get_artificial_defects(self, commit_leftover):
        c, ta = self._COMMIT_HEIGHT, self._CODE.TIME_AXIS
        return {self._CODE.raise_node(v, -c)
                for v in self._CODE.get_syndrome(commit_leftover) if v[ta] == c}

    def _get_leftover(self, error, correction):
        commit = set(self._COMMIT_EDGES)
        return ({e for e in error if e in commit} ^ {e for e in correction if e in commit},
                {e for e in error if e not in commit})

    def run(self, decoder, noise_level: float, n: int, shots: int = 1, **kwargs):
        """``shots`` independent runs of ``n`` steady-state decoding cycles each
        (``_schemes.py:392-450``; the reference runs exactly one).  ``step_counts``
        gets ``decoder.decode``'s return value per timed cycle: ``None`` for UF,
        ``(tSV, tBP)`` for Macar."""
        if n < 1:
            raise ValueError("n must be positive integer.")
        self.reset()
        m, steps = decoder._mc_forward(noise_level, n, shots=shots, **kwargs)
        returns_steps = getattr(decoder, '_DECODER_ID', 0) != 0
        self.step_counts.extend((int(a), int(b)) if returns_steps else None for a, b in steps)
        return m, shots * (self._BUFFER_HEIGHT + (n + 1) * self._COMMIT_HEIGHT) / self._CODE.D


class Frugal(_Streaming):
    """One layer in, one layer committed per decoding cycle (``_schemes.py:480-649``)."""

    def __str__(self) -> str:
        return 'frugal'

    def __init__(self, code, commit_height, buffer_height, commit_edges):
        super().__init__(code, commit_height, buffer_height, commit_edges)
        self.error: set[Edge] = set()
        self._future_boundary_syndrome: set[Node] = set()

    def reset(self):
        super().reset()
        self.error.clear()
        self._future_boundary_syndrome.clear()

    def get_logical_error(self) -> int:
        return self._count_and_lower()

    def run(self, decoder, noise_level: float, n: int, time_only='merging', shots: int = 1, **kwargs):
        """``shots`` independent memory experiments of ``n*d`` timed decoding cycles each
        (``_schemes.py:566-649``; the reference runs exactly one, ``shots=1``).  Returns
        ``(logical errors, slenderness)`` summed over the experiments; ``step_counts`` gets
        ``n`` entries per experiment."""
        if n < 1:
            raise ValueError("n must be positive integer.")
        self.reset()
        decoder.reset()
        m, steps = decoder._mc_frugal(noise_level, n, time_only=time_only, shots=shots, **kwargs)
        self.step_counts.extend(steps)
        return m, shots * (math.ceil(self.WINDOW_HEIGHT / self._COMMIT_HEIGHT) / self._CODE.D + n + 1)

    def sample_latency(self, decoder, noise_level: float, draw=False, log_history=False,
                       time_only: str = 'merging', shots=None, seed=None, **kwargs_for_draw_decode):
        """The number of decoder timesteps between the last measurement round and the final
        correction (``_schemes.py:651-698``): a memory experiment of ``WINDOW_HEIGHT + 1`` noisy
        cycles, one more without future-boundary faults, then ``WINDOW_HEIGHT - 1`` noiseless
        ones during which the decoder stops growing and merging once no defect is left; the
        latency is the runtime of the last ``WINDOW_HEIGHT`` cycles.  Assumes commit height 1.

        Returns one sample like the reference; ``shots=k`` returns a list of ``k`` independent
        samples from one kernel launch (what ``sim.latency.frugal`` uses)."""
        if draw or log_history:
            raise NotImplementedError("Drawing / history are outside this engine.")
        if time_only not in ('all', 'merging', 'unrooting'):
            raise ValueError(f'Unknown time_only: {time_only}.')
        self.reset()
        decoder.reset()
        lat = decoder._mc_latency(noise_level, 1 if shots is None else shots, seed=seed)
        column = lat[:, ('all', 'merging', 'unrooting').index(time_only)]
        return int(column[0]) if shots is None else [int(x) for x in column]
