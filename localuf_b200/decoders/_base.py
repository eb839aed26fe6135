"""Common host-side behaviour of the device-backed decoders.

Mirrors ``localuf/_base_classes.py:641-712`` (``Decoder``) and the UF-family
state of ``localuf/decoders/_base_uf.py:33-99`` (``growth``, ``erasure``,
``syndrome``): the attributes exist with the reference's names and types, but
they are views of bit arrays that a CUDA kernel produced.
"""
from __future__ import annotations

import random

import numpy as np

from localuf_b200 import _dcs, _engine, dist
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
            seed = dist.allreduce_counts([seed if rank == 0 else 0], device=f'cuda:{self.engine.device}')[0]   # rank 0's seed for everybody
        lo, cnt = dist.shard(n, rank, size)
        fails, shots = run_local(seed, lo, cnt) if cnt else (0, 0)
        fails, shots = dist.allreduce_counts([fails, shots], device=f'cuda:{self.engine.device}')
        assert shots == n
        return fails


    def _mc_weight(self, weight, n: int, seed=None) -> int:
        """Failure count of ``n`` shots of weight-``weight`` errors (``Batch.sim_cycles_given_weight``)."""
        eng = self.engine
        if isinstance(weight, (int, np.integer)):
            if not 0 <= weight <= len(self._CODE.NOISE.FRESH_EDGES):
                raise ValueError('Sample larger than population or is negative')
            return self._sharded_counts(
                lambda s, lo, cnt: eng.mc_weight(self._DECODER_ID, self._flags, int(weight), s, lo, cnt), n, seed)
        # circuit-level: weight vectors over fault populations; the reference's sampler, in batches
        flat, noise = self._CODE.FLAT, self._CODE.NOISE
        rank, size = dist.world()
        lo, cnt = dist.shard(n, rank, size)
        # the forcers draw from Python's global `random` stream like the reference (noise/forcers.py).  One rank:
        # that stream as it stands (or seeded with `seed`).  Several ranks: every rank would draw the SAME errors
        # from equally seeded streams and the all-reduced count would hold each shot `size` times, so each rank
        # draws from its own stream derived from rank 0's seed, and the caller's stream is put back afterwards.
        state = None
        if size > 1 or seed is not None:
            seed = self._draw_seed(seed)
            if size > 1:
                seed = dist.allreduce_counts([seed if rank == 0 else 0], device=f'cuda:{eng.device}')[0]
            state = random.getstate()
            random.seed(seed * 1_000_003 + rank)
        fails, done = 0, 0
        try:
            while done < cnt:
                k = min(4096, cnt - done)
                errors = [noise.force_error(weight) for _ in range(k)]
                err = np.stack([flat.edges_to_bits(e) for e in errors])
                syn = np.stack([flat.syndrome_to_bits(self._CODE.get_syndrome(e)) for e in errors])
                fails += eng.mc_given_errors(self._DECODER_ID, self._flags, err, syn)[0]
                done += k
        finally:
            if state is not None:
                random.setstate(state)
        return dist.allreduce_counts([fails], device=f'cuda:{eng.device}')[0]

    def subset_sample(self, noise_level: float, n: int, tol: float = 5e-1, **kwargs):
        """Failures per error subset (``_base_classes.py:682-712``): a DataFrame indexed by
        weight with columns ``['subset prob', 'survival prob', 'm', 'n']``, subsets taken in
        decreasing probability until the probability left over is below ``tol`` times the
        failure rate accumulated so far."""
        import pandas as pd
        subset_probs = self._CODE.NOISE.subset_probabilities(noise_level)
        mean, mn, weights = 0, [], []
        for weight in subset_probs.index:
            weights.append(weight)
            weightless = weight == 0 if isinstance(weight, (int, np.integer)) else all(w == 0 for w in weight)
            if weightless:
                mn.append((np.nan, np.nan))
            else:
                w = int(weight) if isinstance(weight, (int, np.integer)) else tuple(int(x) for x in weight)
                m, shots = self._CODE.SCHEME.sim_cycles_given_weight(self, w, n, **kwargs)
                mn.append((m, shots))
                mean += subset_probs.loc[weight, 'subset prob'] * m / shots
                if mean * tol > subset_probs.loc[weight, 'survival prob']:
                    break
        index = pd.MultiIndex.from_tuples(weights) if type(subset_probs.index) is pd.MultiIndex else pd.Index(weights)
        subset_failures = pd.DataFrame(mn, index=index, columns=['m', 'n'])
        return pd.concat([subset_probs, subset_failures], axis=1, join='inner')

    def _mc_global(self, noise_level: float, shots: int = 1, seed=None):
        """``shots`` windows of the global batch scheme (``_schemes.py:186-203``): sampled and
        decoded on the device, the logical error strings of each leftover counted there too
        (``luf_count_strings``: the in-kernel ``Pairs`` of the forward scheme, ascending edge id)."""
        eng = self.engine
        rank, size = dist.world()
        seed = self._draw_seed(seed)
        if size > 1:
            seed = dist.allreduce_counts([seed if rank == 0 else 0], device=f'cuda:{eng.device}')[0]
        lo, cnt = dist.shard(shots, rank, size)
        total = 0
        if cnt:           # sample -> decode -> count strings, all on the device (luf_count_strings)
            total = eng.mc_global(self._DECODER_ID, self._flags, noise_level, seed, lo, cnt)
        return dist.allreduce_counts([total], device=f'cuda:{eng.device}')[0]

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
            seed = dist.allreduce_counts([seed if rank == 0 else 0], device=f'cuda:{eng.device}')[0]
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

    # ---- decoder confidence scores (``_base_uf.py:101-121,250-316``; kernel ``csrc/luf_dcs.cuh``)
    def _dcs_engine(self, noise_level) -> '_dcs.DcsEngine':
        """The swim graph with the edge weights of ``noise_level``, uploaded once per level."""
        cache = self.__dict__.setdefault('_dcs_engines', {})
        if noise_level not in cache:
            cache[noise_level] = _dcs.DcsEngine(_dcs.flat_tables(self._CODE, noise_level), device=self.engine.device)
        return cache[noise_level]

    def unclustered_edge_fraction(self, noise_level=None) -> float:
        """``_base_uf.py:106-121``: the weighted fraction of edge growth that is *not* in a cluster."""
        _, uef, _ = self._dcs_engine(noise_level).run_host(_dcs.pack_growth(self._growth_codes[None, :]),
                                                           want_swim=False, want_clustered=False)
        return float(uef[0])

    def swim_distance(self, noise_level=None, draw=False, **kwargs_for_draw_swim_graph) -> float:
        """``_base_uf.py:250-279``: the shortest west-east distance when travelling inside clusters is free.
        Assumes ``validate`` or ``decode`` has been called."""
        if draw:
            raise NotImplementedError("Drawing is outside this engine (SURVEY.md section 2, row 14).")
        swim, _, _ = self._dcs_engine(noise_level).run_host(_dcs.pack_growth(self._growth_codes[None, :]),
                                                            want_uef=False, want_clustered=False)
        return float(swim[0])

    def sample_confidence_scores(self, noise_level: float, n: int, noise_level_for_priors=None, seed=None,
                                 chunk: int = 1 << 16):
        """``n`` sampled shots, decoded and scored on the device without leaving it: sample (K1) -> decode
        with growth export (K2 / K3a) -> logical parity (K4) -> confidence scores (``k_dcs``).  Returns a
        dict of arrays ``swim_distance``, ``unclustered_edge_fraction``, ``unclustered_node_fraction`` and
        ``logical`` (uint8), one entry per shot -- the batched form of calling ``decode`` and then the three
        score methods ``n`` times."""
        import torch
        eng, dcs = self.engine, self._dcs_engine(noise_level_for_priors)
        dev = torch.device('cuda', eng.device)
        seed = self._draw_seed(seed)
        out = dict(swim_distance=np.zeros(n), unclustered_edge_fraction=np.zeros(n),
                   unclustered_node_fraction=np.zeros(n), logical=np.zeros(n, np.uint8))
        with torch.cuda.device(dev):
            k = min(chunk, max(n, 1))
            err = torch.empty((k, eng.WE), dtype=torch.int32, device=dev)
            syn = torch.empty((k, eng.WD), dtype=torch.int32, device=dev)
            corr = torch.empty((k, eng.WE), dtype=torch.int32, device=dev)
            growth = torch.empty((k, eng.WG), dtype=torch.int32, device=dev)
            logical = torch.empty(k, dtype=torch.uint8, device=dev)
            fails = torch.zeros(1, dtype=torch.int64, device=dev)
            swim = torch.empty(k, dtype=torch.float64, device=dev)
            uef = torch.empty(k, dtype=torch.float64, device=dev)
            cl = torch.empty(k, dtype=torch.int32, device=dev)
            done = 0
            while done < n:
                c = min(k, n - done)
                eng.sample(noise_level, seed, done, c, err, syn)
                eng.decode_batch(self._DECODER_ID, self._flags, c, syn, corr, growth2b=growth)
                eng.check(c, err, corr, logical, fails)
                dcs.run(c, growth, eng.WG, swim=swim, uef=uef, clustered=cl)
                sl = slice(done, done + c)
                out['swim_distance'][sl] = swim[:c].cpu().numpy()
                out['unclustered_edge_fraction'][sl] = uef[:c].cpu().numpy()
                out['unclustered_node_fraction'][sl] = 1 - cl[:c].cpu().numpy() / self._CODE.NODE_COUNT
                out['logical'][sl] = logical[:c].cpu().numpy()
                done += c
        return out

    def _decode_one(self, syndrome):
        """Decode one syndrome on the device; returns (correction set, steps row)."""
        flat = self._CODE.FLAT
        syn = flat.syndrome_to_bits(syndrome)[None, :]
        corr, growth2b, steps = self.engine.decode_batch_host(
            self._DECODER_ID, self._flags, syn, want_growth=True, want_steps=True)
        self._growth_codes = _engine.unpack_growth(growth2b, flat.n_edges)[0]
        return flat.bits_to_edges(corr[0]), steps[0]


__all__ = ['Decoder', 'BaseUF', 'Growth']
