"""Noise models: which edges flip, and with what probability.

Mirrors ``localuf/noise/main.py`` (``_Uniform`` ``:92-149``, ``CodeCapacity`` /
``Phenomenological`` ``:152-169``, ``CircuitLevel`` ``:172-362``) and
``localuf/noise/_multiset_handler.py:9-19``.

Two consumers:

* the drop-in host API (``Code.make_error``) draws from Python's global
  ``random`` stream in the reference's edge order, so seeded error sets are
  identical to the reference's;
* the device sampler (kernel K1) only needs *probability classes*: every edge
  belongs to one class, every class has one flip probability at a given noise
  level (``classes()`` / ``class_probabilities()``).
"""
from __future__ import annotations

import abc
import random
from functools import cached_property, lru_cache
from math import log10, prod

Edge = tuple

_DEFAULT_MULTIPLICITIES = {
    # edge family -> (#two-qubit-gate faults of the 4/15 kind, of the 8/15 kind,
    #                 #prep/reset faults, #measurement faults); noise/main.py:175-188
    'S': (4, 2, 1, 0),
    'E westmost': (1, 0, 2, 0),
    'E bulk': (2, 0, 1, 0),
    'E eastmost': (1, 0, 1, 0),
    'U 3': (3, 0, 1, 1),
    'U 4': (4, 0, 0, 1),
    'SD': (2, 0, 0, 0),
    'EU west corners': (2, 0, 1, 0),
    'EU east corners': (2, 0, 2, 0),
    'EU edge': (3, 0, 1, 0),
    'EU centre': (4, 0, 0, 0),
    'SEU': (2, 0, 0, 0),
}
_HORIZONTALS = ('S', 'E westmost', 'E bulk', 'E eastmost')
_VERTICALS = ('U 3', 'U 4')
_COEFFICIENTS = {
    # pi = c * p; noise/main.py:191-195
    'standard': (4 / 15, 8 / 15, 2 / 3, 1),
    'balanced': (4 / 15, 8 / 15, 8 / 15, 8 / 15),
    'ion trap': (4 / 15, 8 / 15, (2e-3) / 3, 1e-2),
}


def odd_fault_probability(multiplicities, pi) -> float:
    """Pr(odd number of faults) for a multiset of independent faults, to first
    order as the reference computes it (``_multiset_handler.py:17-19``)."""
    none = prod((1 - p) ** m for p, m in zip(pi, multiplicities))
    one = sum(m * p / (1 - p) for p, m in zip(pi, multiplicities))
    return none * one


class Noise(abc.ABC):
    @abc.abstractmethod
    def __str__(self) -> str: ...

    @abc.abstractmethod
    def make_error(self, noise_level: float) -> set[Edge]: ...

    @abc.abstractmethod
    def classes(self) -> list[tuple[Edge, ...]]:
        """Fresh edges grouped into probability classes (sampling order)."""

    @abc.abstractmethod
    def class_probabilities(self, noise_level: float) -> list[float]:
        """Flip probability of each class of :meth:`classes` at ``noise_level``."""

    @abc.abstractmethod
    def get_edge_weights(self, noise_level): ...

    @staticmethod
    def log_odds_of_no_flip(flip_probability: float) -> float:
        return log10((1 - flip_probability) / flip_probability)

    @abc.abstractmethod
    def force_error(self, weight):
        """Make an error of weight ``weight`` (``noise/main.py:38-40``)."""

    @abc.abstractmethod
    def subset_probability(self, weights, noise_level: float):
        """Probability of any error of weight ``weight``, for ``weight`` in ``weights``."""

    def subset_probabilities(self, noise_level: float, survival: bool = True):
        """DataFrame indexed by subset weight with columns ``['subset prob', 'survival prob']``,
        most probable subset first (``noise/main.py:50-66``)."""
        import pandas as pd
        subset_prob = self.subset_probability(self.ALL_WEIGHTS, noise_level)
        df = pd.DataFrame({'subset prob': subset_prob}, index=self.ALL_WEIGHTS_INDEX)
        if survival:
            df.sort_values(by='subset prob', inplace=True)
            df['survival prob'] = df['subset prob'].cumsum().shift(fill_value=0)
        df.sort_values(by='subset prob', ascending=False, inplace=True)
        return df


class _Uniform(Noise):
    """Every fresh edge flips with probability ``noise_level``."""

    def __init__(self, fresh_edges, all_edges=None):
        self._FRESH_EDGES = tuple(fresh_edges)
        self._ALL_EDGES = self._FRESH_EDGES if all_edges is None else tuple(all_edges)

    @property
    def FRESH_EDGES(self):
        return self._FRESH_EDGES

    def make_error(self, noise_level):
        rnd = random.random
        return {e for e in self._FRESH_EDGES if rnd() < noise_level}

    def force_error(self, weight: int):
        return set(random.sample(self._FRESH_EDGES, weight))

    @property
    def ALL_WEIGHTS(self):
        return range(len(self._FRESH_EDGES) + 1)

    @property
    def ALL_WEIGHTS_INDEX(self):
        import pandas as pd
        return pd.Index(self.ALL_WEIGHTS, name='weight')

    def subset_probability(self, weights, noise_level: float):
        """Binomial weight distribution (``noise/main.py:128-134``)."""
        from scipy.stats import binom
        return binom.pmf(k=weights, n=len(self._FRESH_EDGES), p=noise_level)

    def classes(self):
        return [self._FRESH_EDGES]

    def class_probabilities(self, noise_level):
        return [float(noise_level)]

    def get_edge_weights(self, noise_level):
        if noise_level is None:
            pair = (0, 1)
        else:
            pair = (noise_level, self.log_odds_of_no_flip(noise_level))
        return {e: pair for e in self._ALL_EDGES}


class CodeCapacity(_Uniform):
    def __str__(self): return 'code capacity'


class Phenomenological(_Uniform):
    def __str__(self): return 'phenomenological'


class _BaseForceByPair:
    """Error subsets distinguished by how many faults of each pair type occurred
    (``noise/forcers.py:76-128``): population ``k`` lists every edge once per fault of type
    ``k`` it stands for; an edge drawn twice flips back."""

    def __init__(self, edges):
        self._EDGES = edges

    def _populations(self):
        out = []
        for k in range(4):
            pairs = []
            for m, edges in self._EDGES.items():
                pairs += m[k] * list(edges)
            out.append(pairs)
        return tuple(out)

    def force_error(self, weight):
        error: set[Edge] = set()
        if len(weight) != len(self.PAIR_POPULATIONS):
            raise ValueError('zip() arguments have different lengths')
        for pairs, w in zip(self.PAIR_POPULATIONS, weight):
            for e in random.sample(pairs, w):
                error.symmetric_difference_update((e,))
        return error

    @property
    def ALL_WEIGHTS(self):
        import itertools
        return tuple(itertools.product(*(range(len(pairs) + 1) for pairs in self.PAIR_POPULATIONS)))

    def subset_probability(self, weights, pi):
        import numpy as np
        from scipy.stats import binom
        probs = np.ones(len(weights))
        columns = np.array(weights).T
        for row, pairs, p in zip(columns, self.PAIR_POPULATIONS, pi):
            probs *= binom.pmf(k=row, n=len(pairs), p=p)
        return probs


class ForceByPair(_BaseForceByPair):
    """Weights are 4-tuples (``noise/forcers.py:131-141``); any parametrization but 'balanced'."""

    @cached_property
    def PAIR_POPULATIONS(self):
        return self._populations()


class ForceByPairBalanced(_BaseForceByPair):
    """The last three pair types share one probability: weights are 2-tuples
    (``noise/forcers.py:144-155``)."""

    @cached_property
    def PAIR_POPULATIONS(self):
        p0, *rest = self._populations()
        return (p0, sum(rest, start=[]))


class ForceByEdge:
    """Weights count the flipped edges of each multiplicity class (``noise/forcers.py:158-184``)."""

    def __init__(self, edges):
        self._EDGES = edges

    def force_error(self, weight):
        if len(weight) != len(self._EDGES):
            raise ValueError(f'len(weight)={len(weight)} != {len(self._EDGES)}=number of edge subsets')
        error: set[Edge] = set()
        for subset, w in zip(self._EDGES.values(), weight):
            error.update(random.sample(subset, w))
        return error

    @property
    def ALL_WEIGHTS(self):
        import itertools
        return tuple(itertools.product(*(range(len(subset) + 1) for subset in self._EDGES.values())))

    def subset_probability(self, weights, pi):
        raise NotImplementedError


class CircuitLevel(Noise):
    """Circuit-level depolarising noise: each edge flips with the probability
    that an odd number of the circuit faults it stands for occurred."""

    def __str__(self): return 'circuit-level'

    def __init__(self, fresh_edge_dict, parametrization, demolition, monolingual,
                 fresh_merges=None, force_by='pair', all_edge_dict=None, all_merges=None):
        self._FRESH_EDGES = self._group(fresh_edge_dict, demolition, monolingual, fresh_merges)
        self._COEFFICIENTS = _COEFFICIENTS[parametrization]
        self._ALL_EDGES = (self._FRESH_EDGES if all_edge_dict is None
                           else self._group(all_edge_dict, demolition, monolingual, all_merges))
        self._probs = lru_cache(maxsize=None)(self._compute_probabilities)
        if force_by == 'pair':                      # noise/main.py:219-226
            self._FORCER = (ForceByPairBalanced if parametrization == 'balanced' else ForceByPair)(self._FRESH_EDGES)
        elif force_by == 'edge':
            self._FORCER = ForceByEdge(self._FRESH_EDGES)
        else:
            raise ValueError(f"Unknown force_by {force_by!r}.")

    @staticmethod
    def _make_multiplicities(demolition: bool, monolingual: bool):
        """Per-family multiplicity vectors (``noise/main.py:262-288``)."""
        out = {k: list(v) for k, v in _DEFAULT_MULTIPLICITIES.items()}
        if demolition:
            for k in _HORIZONTALS:
                out[k][2] += 1
            for k in _VERTICALS:
                out[k][3] += 1
        if monolingual:
            for k in _HORIZONTALS + _VERTICALS:
                out[k][2] += 2
        return {k: tuple(v) for k, v in out.items()}

    @classmethod
    def _group(cls, edge_dict, demolition, monolingual, merges):
        """multiplicity vector -> tuple of edges, in the reference's first-seen
        order (``noise/main.py:240-307``); a redundant edge adds its vector to
        its substitute's."""
        mult = cls._make_multiplicities(demolition, monolingual)
        grouped: dict[tuple, list[Edge]] = {}
        if merges is None:
            for family, edges in edge_dict.items():
                grouped.setdefault(mult[family], []).extend(edges)
        else:
            total: dict[Edge, list[int]] = {}
            for family, edges in edge_dict.items():
                m = mult[family]
                for e in edges:
                    acc = total.setdefault(merges.get(e, e), [0, 0, 0, 0])
                    for k in range(4):
                        acc[k] += m[k]
            for e, acc in total.items():
                grouped.setdefault(tuple(acc), []).append(e)
        return {m: tuple(es) for m, es in grouped.items()}

    def _pi(self, noise_level):
        return tuple(c * noise_level for c in self._COEFFICIENTS)

    def _compute_probabilities(self, noise_level):
        pi = self._pi(noise_level)
        return {m: odd_fault_probability(m, pi) for m in self._FRESH_EDGES}

    def _get_flip_probabilities(self, noise_level):
        return self._probs(noise_level)

    def make_error(self, noise_level):
        rnd = random.random
        error: set[Edge] = set()
        for m, pr in self._probs(noise_level).items():
            error.update(e for e in self._FRESH_EDGES[m] if rnd() < pr)
        return error

    def force_error(self, weight):
        return self._FORCER.force_error(weight)

    @property
    def ALL_WEIGHTS(self):
        return self._FORCER.ALL_WEIGHTS

    @property
    def ALL_WEIGHTS_INDEX(self):
        import pandas as pd
        return pd.MultiIndex.from_tuples(self.ALL_WEIGHTS)

    def subset_probability(self, weights, noise_level: float):
        return self._FORCER.subset_probability(weights, self._pi(noise_level))

    def classes(self):
        return list(self._FRESH_EDGES.values())

    def class_probabilities(self, noise_level):
        return [float(x) for x in self._probs(noise_level).values()]

    def get_edge_weights(self, noise_level):
        out = {}
        if noise_level is None:
            for edges in self._ALL_EDGES.values():
                for e in edges:
                    out[e] = (0, 1)
            return out
        pi = self._pi(noise_level)
        for m, edges in self._ALL_EDGES.items():
            pr = odd_fault_probability(m, pi)
            w = self.log_odds_of_no_flip(pr)
            for e in edges:
                out[e] = (pr, w)
        return out
