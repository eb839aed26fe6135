"""Snowflake: the streaming local Union-Find decoder, device-backed.

Same constructor and compatibility errors as the reference
(``localuf/decoders/snowflake/main.py:41-78``).  The decoder only exists
together with the frugal scheme: ``code.SCHEME.run(decoder, p, n)``
(``_schemes.py:566-649``) runs whole memory experiments -- window raise, noise,
drop, grow, merge, floor commit, string pairing, logical count -- inside
``csrc/luf_sf.cuh`` with the window resident in shared memory.
"""
from __future__ import annotations

from collections import Counter, defaultdict

import numpy as np

from localuf_b200 import _dcs, _engine, _window, dist
from localuf_b200.constants import GROWTH_FROM_DEVICE
from localuf_b200.codes import Repetition
from localuf_b200.decoders._base import Decoder
from localuf_b200.decoders._bitstrings import Bitstrings
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
        # bit-string I/O (snowflake/main.py:104-123): the floor edges in output order, as local
        # indices of the bottom layer block
        self._BITSTRING_CONVERTER = Bitstrings(code.D, code.SCHEME.WINDOW_HEIGHT, type(code) is Repetition)
        index = {e: k for k, e in enumerate(code.EDGES)}
        self._LOWEST_EDGES = self._BITSTRING_CONVERTER.lowest_edges(
            str(code.NOISE), _include_timelike_lowest_edges=_include_timelike_lowest_edges)
        self._lowest_local = np.array(
            [self._tables.layered_of_code[index[e]] % self._tables.EL for e in self._LOWEST_EDGES], np.int64)
        self._cycles: list = []            # syndrome rows of the decoding cycles since reset()
        self._drops: list = []             # per cycle: the decoder only dropped (defects_possible=False)
        self._state = None                 # (number of cycles, growth export, mins) of the last replay with metrics
        self._dcs_engines: dict = {}
        self.confidence_score_history = defaultdict(list)      # snowflake/main.py:124-126
        self.min_active_layers: Counter = Counter()            # snowflake/main.py:127-128

    def __repr__(self) -> str:
        return f'decoders.Snowflake({self.CODE})'

    @property
    def engine(self) -> '_engine.StreamEngine':
        if self._stream_engine is None:
            self._stream_engine = _engine.StreamEngine(self._CODE, self._tables)
        return self._stream_engine

    # -- single decoding cycles and bit-string I/O ------------------------------
    def reset(self):
        """``snowflake/main.py:201-214``: an empty window; forgets ``floor_history``."""
        super().reset()
        self._cycles = []
        self._drops = []
        self._state = None
        self.confidence_score_history = defaultdict(list)
        self.min_active_layers = Counter()
        if hasattr(self, 'floor_history'):
            del self.floor_history

    def init_history(self):
        """``snowflake/main.py:181-184`` (only ``floor_history`` is kept here)."""
        self.floor_history: list = []

    def _syndrome_row(self, syndrome) -> np.ndarray:
        t = self._tables
        row = np.zeros((t.EL + 31) // 32, np.uint32)
        top = t.h - 1
        for v in syndrome:
            if v[-1] != top:
                raise ValueError(f'Defect {v} is not in the top layer {top} of the viewing window.')
            q = t.position_of(v)
            if q < t.NLb:
                raise ValueError(f'{v} is a boundary node, not a detector.')
            row[q >> 5] ^= np.uint32(1 << (q & 31))
        return row

    def _floor_vector(self, floor_row: np.ndarray) -> str:
        bits = (floor_row[self._lowest_local >> 5] >> (self._lowest_local & 31).astype(np.uint32)) & 1
        return ''.join('1' if b else '0' for b in bits)

    def _run_cycles(self, rows, drops=None):
        """One window, the given syndrome rows in order -> (runtimes [n, 2], floor rows [n, W]).
        ``drops``: per cycle, whether the decoder only drops (``defects_possible=False``)."""
        if drops is not None and any(drops):
            rt, floor, _, _, _ = self.engine.decode_metrics(self._flags | _engine.SF_SYNDROME, np.stack(rows)[None],
                                                            drop_only=np.array(drops, np.uint8)[None])
            return rt[0], floor[0]
        rt, floor, _ = self.engine.decode_host(self._flags | _engine.SF_SYNDROME, np.stack(rows)[None])
        return rt[0], floor[0]

    def decode(self, syndrome, log_history=False, log_floor_history: bool = False, metrics=(),
               noise_level_for_priors=None, time_only: str = 'merging', defects_possible: bool = True):
        """One decoding cycle (``snowflake/main.py:216-281``): ``syndrome`` = the defects of the
        newly exposed top layer.  Returns the number of timesteps of the cycle.

        The device keeps a window inside one kernel launch, so the state between calls is the
        list of syndromes since :meth:`reset`; each call replays it (cost grows with the number
        of cycles -- use :meth:`generate_output` or ``Frugal.run`` for long streams)."""
        if log_history:
            raise NotImplementedError("Drawing / history are outside this engine.")
        if time_only not in ('all', 'merging', 'unrooting'):
            raise ValueError(f'Unknown time_only: {time_only}.')
        metrics = tuple(metrics)
        self._check_metrics(metrics)
        self._cycles.append(self._syndrome_row(syndrome))
        self._drops.append(not defects_possible)       # the decoder only drops (snowflake/main.py:244-264)
        self._state = None
        if metrics:
            rt, floor, growth, mins = self._replay_with_metrics()
        else:
            rt, floor = self._run_cycles(self._cycles, self._drops)
        if log_floor_history:
            self.floor_history.append(self._floor_vector(floor[-1]))
        runtime = self._runtime(rt[-1], time_only) if defects_possible else (1 if time_only == 'all' else 0)
        if metrics:
            swim = uef = None
            if 'swim_distance' in metrics or 'unclustered_edge_fraction' in metrics:
                swim, uef, _ = self._scores(growth[-1:], noise_level_for_priors)
            self._record_metrics(metrics, [runtime], mins[-1:], swim, uef)
        return runtime

    # ---- decoder confidence scores (snowflake/main.py:266-297; _base_uf.py:101-121,250-316)
    _METRICS = ('throughput', 'swim_distance', 'min_defect_height', 'unclustered_edge_fraction', 'min_active_layer')

    def _check_metrics(self, metrics):
        for metric in metrics:
            if metric not in self._METRICS:
                raise ValueError(f'Unknown metric: {metric}.')

    def _dcs_engine(self, noise_level) -> '_dcs.DcsEngine':
        if noise_level not in self._dcs_engines:
            self._dcs_engines[noise_level] = _dcs.DcsEngine(
                _dcs.window_tables(self._CODE, self._tables, noise_level), device=self.engine.device)
        return self._dcs_engines[noise_level]

    def _scores(self, growth_dev, noise_level):
        """(swim distance, unclustered edge fraction) of growth rows that are already on the device."""
        import torch
        n = growth_dev.shape[0]
        swim = torch.empty(n, dtype=torch.float64, device=growth_dev.device)
        uef = torch.empty(n, dtype=torch.float64, device=growth_dev.device)
        with torch.cuda.device(growth_dev.device):
            self._dcs_engine(noise_level).run(n, growth_dev, growth_dev.shape[1], swim=swim, uef=uef)
        return swim.cpu().numpy(), uef.cpu().numpy(), None

    def _record_metrics(self, metrics, runtimes, mins, swim, uef):
        """``snowflake/main.py:266-281`` for a run of cycles (one entry per cycle and metric)."""
        for k, runtime in enumerate(runtimes):
            for metric in metrics:
                if metric == 'throughput':
                    value = 1 / runtime
                elif metric == 'swim_distance':
                    value = float(swim[k])
                elif metric == 'min_defect_height':
                    value = int(mins[k][0])
                elif metric == 'unclustered_edge_fraction':
                    value = float(uef[k])
                else:
                    self.min_active_layers[int(mins[k][1])] += 1
                    continue
                self.confidence_score_history[metric].append(value)

    def _replay_with_metrics(self):
        """Replay the cycles since ``reset`` with the per-cycle export; the last replay is kept."""
        if self._state is None or self._state[0] != len(self._cycles):
            rows = self._cycles if self._cycles else [np.zeros(self.engine.W, np.uint32)]
            drops = np.array(self._drops if self._cycles else [False], np.uint8)[None]
            rt, floor, _, growth, mins = self.engine.decode_metrics(self._flags | _engine.SF_SYNDROME, np.stack(rows)[None],
                                                                    drop_only=drops if drops.any() else None)
            self._state = (len(self._cycles), rt[0], floor[0], growth, mins[0])
        return self._state[1:]

    def _current(self):
        if not self._cycles:
            raise RuntimeError('no decoding cycle since reset(): the window is empty')
        return self._replay_with_metrics()

    def swim_distance(self, noise_level=None, draw=False, **kwargs) -> float:
        """``_base_uf.py:250-279`` on the window after the last decoding cycle."""
        if draw:
            raise NotImplementedError("Drawing is outside this engine.")
        _, _, growth, _ = self._current()
        return float(self._scores(growth[-1:], noise_level)[0][0])

    def unclustered_edge_fraction(self, noise_level=None) -> float:
        """``_base_uf.py:106-121`` on the window after the last decoding cycle."""
        _, _, growth, _ = self._current()
        return float(self._scores(growth[-1:], noise_level)[1][0])

    def min_defect_height(self) -> int:
        """``snowflake/main.py:283-288``."""
        return int(self._current()[3][-1][0])

    def min_active_layer(self) -> int:
        """``snowflake/main.py:290-296``."""
        return int(self._current()[3][-1][1])

    @property
    def growth(self) -> dict:
        """``snowflake/main.py:151-153``: edge -> Growth over the edges below the future boundary."""
        t, edges = self._tables, self._CODE.EDGES
        ELp = 32 * ((t.EL + 31) // 32)
        if self._cycles:
            codes = _engine.unpack_growth(self._current()[2][-1:].cpu().numpy().view(np.uint32), t.h * ELp)[0]
        else:
            codes = np.zeros(t.h * ELp, np.uint8)
        ta, h = self._CODE.TIME_AXIS, t.h
        out = {}
        for k, e in enumerate(edges):
            if all(v[ta] < h for v in e):
                b, l = divmod(int(t.layered_of_code[k]), t.EL)
                out[e] = GROWTH_FROM_DEVICE[int(codes[b * ELp + l])]
        return out

    def _runtime(self, rt, time_only: str) -> int:
        if time_only == 'all':
            return int(rt[0])
        if time_only == 'merging':
            return int(rt[0]) - 2 * self._LOOPS
        return int(rt[1])

    def generate_output(self, syndromes, output_to_csv_file=None, draw=False, **_kwargs):
        """``snowflake/main.py:656-726``: syndrome vectors (strings of '0' / '1') or sets of
        top-layer defects, one per decoding cycle, in; the floor vectors -- the correction bits
        of ``self._LOWEST_EDGES`` just before each drop -- out, padded with ``WINDOW_HEIGHT``
        empty cycles so that every correction leaves the window.  Optionally saved as CSV
        (``syndrome_vector``, ``floor_history``)."""
        if draw:
            raise NotImplementedError("Drawing is outside this engine.")
        conv, h = self._BITSTRING_CONVERTER, self._tables.h
        first, *_ = syndromes
        if type(first) is str:
            vectors = list(syndromes) + h * ['0' * conv.LENGTH]
            sets = conv.syndrome_vectors_to_sets(vectors)
        else:
            sets = list(syndromes) + h * [set()]
            vectors = conv.sets_to_syndrome_vectors(sets)
        self.init_history()
        _, floor = self._run_cycles(self._cycles + [self._syndrome_row(s) for s in sets], self._drops + [False] * len(sets))
        self._cycles += [self._syndrome_row(s) for s in sets]
        self._drops += [False] * len(sets)
        self.floor_history += [self._floor_vector(f) for f in floor[len(floor) - len(sets):]]
        if output_to_csv_file is not None:
            import csv
            with open(output_to_csv_file, 'w', newline='') as f:
                w = csv.writer(f)
                w.writerow(['syndrome_vector', 'floor_history'])
                w.writerows(zip(vectors, self.floor_history, strict=True))
        return self.floor_history

    def _mc_latency(self, noise_level: float, shots: int, seed=None):
        """``shots`` independent ``Frugal.sample_latency`` experiments -> int32 [shots, 3]
        (time_only = 'all', 'merging', 'unrooting'); under torchrun every rank gets all of them."""
        eng = self.engine
        rank, size = dist.world()
        seed = self._draw_seed(seed)
        if size > 1:
            seed = dist.allreduce_counts([seed if rank == 0 else 0], device=f'cuda:{eng.device}')[0]
        lo, cnt = dist.shard(shots, rank, size)
        out = np.zeros((shots, 3), np.int64)
        if cnt:
            out[lo:lo + cnt] = eng.latency_host(self._flags, noise_level, seed, lo, cnt)
        if size > 1:
            out = np.array(dist.allreduce_counts(out.reshape(-1).tolist(), device=f'cuda:{eng.device}')).reshape(shots, 3)
        return out.astype(np.int32)

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
        if draw or log_history:
            raise NotImplementedError("Drawing / history are outside this engine.")
        metrics = tuple(metrics)
        self._check_metrics(metrics)
        if time_only not in ('all', 'merging', 'unrooting'):
            raise ValueError(f'Unknown time_only: {time_only}.')
        eng = self.engine
        rank, size = dist.world()
        seed = self._draw_seed(seed)
        if size > 1:
            seed = dist.allreduce_counts([seed if rank == 0 else 0], device=f'cuda:{eng.device}')[0]
        lo, cnt = dist.shard(shots, rank, size)
        if cnt and metrics:
            # Frugal.run(metrics=...) (_schemes.py:636-644): the scores of every timed cycle, stream after
            # stream (under torchrun: this rank's streams), with noise_level_for_priors = noise_level
            counts, steps = [0, 0, 0, 0], []
            per_stream = n * self._CODE.D * eng.growth_words * 4
            chunk = max(1, min(cnt, (256 << 20) // max(per_stream, 1)))
            for start in range(0, cnt, chunk):
                k = min(chunk, cnt - start)
                c, st, growth, cycle = eng.mc_metrics(self._flags, noise_level, seed, lo + start, k, n)
                counts = [a + b for a, b in zip(counts, c)]
                steps.append(st)
                swim = uef = None
                if 'swim_distance' in metrics or 'unclustered_edge_fraction' in metrics:
                    swim, uef, _ = self._scores(growth, noise_level)
                cyc = cycle.reshape(-1, 4)
                self._record_metrics(metrics, [self._runtime(r[2:], time_only) for r in cyc], cyc[:, :2], swim, uef)
            steps = np.concatenate(steps)
        elif cnt:
            counts, steps = eng.mc_host(self._flags, noise_level, seed, lo, cnt, n, want_steps=True)
        else:
            counts, steps = [0, 0, 0, 0], None
        m = dist.allreduce_counts([counts[0]], device=f'cuda:{eng.device}')[0]
        out = [] if steps is None else [self.steps_from_sums(s, time_only) for s in steps.reshape(-1, 2)]
        return m, out
