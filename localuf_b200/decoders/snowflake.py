"""Snowflake: the streaming local Union-Find decoder, device-backed.

Same constructor and compatibility errors as the reference
(``localuf/decoders/snowflake/main.py:41-78``).  The decoder only exists
together with the frugal scheme: ``code.SCHEME.run(decoder, p, n)``
(``_schemes.py:566-649``) runs whole memory experiments -- window raise, noise,
drop, grow, merge, floor commit, string pairing, logical count -- inside
``csrc/luf_sf.cuh`` with the window resident in shared memory.
"""
from __future__ import annotations

from localuf_b200 import _engine, _window, dist
from localuf_b200.decoders._base import Decoder
from localuf_b200.schemes import Frugal


class Snowflake(Decoder):
    def __init__(self, code, merger: str = 'fast', schedule: str = '2:1', unrooter: str = 'full',
                 _neighbor_order=None, _include_timelike_lowest_edges: bool = True):
        if str(code.NOISE) == 'code capacity':
            raise ValueError('Snowflake incompatible with code capacity noise model.')
        if not isinstance(code.SCHEME, Frugal):
            raise ValueError('Snowflake only compatible with frugal scheme i.e. `code.SCHEME` must be `Frugal`.')
        if merger not in ('fast', 'slow') or schedule not in ('2:1', '1:1'):
            raise ValueError('unknown merger / schedule')
        if unrooter != 'full':
            raise NotImplementedError(
                "Only the default 'full' unrooter is built: the 'simple' unrooter writes a neighbour's "
                "next_cid inside a timestep (snowflake/main.py:1441-1449) and is order dependent in the reference.")
        super().__init__(code)
        self._flags = (_engine.SF_SLOW_MERGER if merger == 'slow' else 0) | \
                      (_engine.SF_ONE_ONE if schedule == '1:1' else 0)
        self._LOOPS = 1 if schedule == '1:1' else 2
        self._tables = _window.build(code, _neighbor_order)
        self._stream_engine = None

    def __repr__(self) -> str:
        return f'decoders.Snowflake({self.CODE})'

    @property
    def engine(self) -> '_engine.StreamEngine':
        if self._stream_engine is None:
            self._stream_engine = _engine.StreamEngine(self._CODE, self._tables)
        return self._stream_engine

    def decode(self, syndrome, **kwargs):
        raise NotImplementedError(
            "The device keeps Snowflake's window inside one kernel for a whole memory experiment; "
            "use code.SCHEME.run(decoder, p, n) (Frugal.run) or StreamEngine.decode_host for given errors.")

    def steps_from_sums(self, sums, time_only: str):
        """(sum of 'all' runtimes, sum of 'unrooting' runtimes) over d cycles -> Frugal.step_counts entry
        (snowflake/main.py:340,347: 'all' - 'merging' = 2 per merge loop)."""
        d = self._CODE.D
        if time_only == 'all':
            return int(sums[0])
        if time_only == 'merging':
            return int(sums[0]) - 2 * self._LOOPS * d
        if time_only == 'unrooting':
            return int(sums[1])
        raise ValueError(f'Unknown time_only: {time_only}.')

    def _mc_frugal(self, noise_level: float, n: int, time_only: str = 'merging', shots: int = 1, seed=None,
                   draw=False, log_history=False, metrics=(), **kwargs):
        """``shots`` independent ``Frugal.run(n)`` memory experiments; returns (m, step_counts)."""
        if draw or log_history or tuple(metrics):
            raise NotImplementedError("Drawing / history / per-cycle metrics are outside this engine.")
        eng = self.engine
        rank, size = dist.world()
        seed = self._draw_seed(seed)
        if size > 1:
            seed = dist.allreduce_counts([seed if rank == 0 else 0])[0]
        lo, cnt = dist.shard(shots, rank, size)
        if cnt:
            counts, steps = eng.mc_host(self._flags, noise_level, seed, lo, cnt, n, want_steps=True)
        else:
            counts, steps = [0, 0, 0, 0], None
        m = dist.allreduce_counts([counts[0]], device=f'cuda:{eng.device}')[0]
        out = [] if steps is None else [self.steps_from_sums(s, time_only) for s in steps.reshape(-1, 2)]
        return m, out
