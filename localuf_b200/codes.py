"""Decoding graphs of the repetition code and the unrotated surface code.

Host-side graph layer of the B200 engine.  It keeps the reference's public
surface (``localuf/_base_classes.py:23-175`` constructor, ``:264-477``
attributes and helpers; ``localuf/codes.py:19-95`` Repetition, ``:98-324``
Surface) and -- unlike the reference, which only ever works on Python sets of
coordinate tuples -- also lowers every graph to flat integer arrays
(:class:`localuf_b200._graph.FlatGraph`) that the CUDA kernels consume.

What has to match the reference bit for bit:

* the order of ``EDGES`` (edge id = position; every canonical tie-break of
  the UF kernel is expressed in edge ids),
* every edge stored ``(u, v)`` with ``u`` the north/west/down-side endpoint,
* ``index_to_id`` (initial cluster ids of Macar/Actis).

``NODES`` is ordered by node id here (the reference exposes CPython set
order, which carries no meaning).
"""
from __future__ import annotations

import abc
from functools import cached_property
from typing import Iterator

from localuf_b200 import noise as _noise
from localuf_b200 import schemes as _schemes
from localuf_b200._graph import FlatGraph

Node = tuple
Edge = tuple


def _shift(v: Node, axis: int, delta: int) -> Node:
    w = list(v)
    w[axis] += delta
    return tuple(w)


class Code(abc.ABC):
    """Decoding graph G = (V, E) plus its noise model and decoding scheme.

    Constructor signature and error behaviour follow
    ``localuf/_base_classes.py:23-124``.
    """

    _TIME_AXIS = -1
    _LONG_AXIS: int
    _CODE_DIMENSION: int

    def __init__(
            self,
            d: int,
            noise: str,
            scheme: str = 'batch',
            window_height: int | None = None,
            commit_height: int | None = None,
            buffer_height: int | None = None,
            merge_equivalent_boundary_nodes: bool = False,
            parametrization: str = 'balanced',
            demolition: bool = False,
            monolingual: bool = False,
            _merge_redundant_edges: bool = True,
    ):
        self._D = d
        self._MERGED_EQUIVALENT_BOUNDARY_NODES = merge_equivalent_boundary_nodes
        cl_kw = dict(parametrization=parametrization, demolition=demolition, monolingual=monolingual)

        if 'batch' in scheme:
            if commit_height is not None or buffer_height is not None:
                height = commit_height if commit_height is not None else buffer_height
                raise ValueError(f"Cannot specify `{height=}` for `{scheme}` scheme.")
            if noise == 'code capacity':
                if window_height is not None:
                    raise ValueError("Cannot specify `window_height` for code capacity noise model.")
                h = 1
                self._N_EDGES, self._EDGES = self._code_capacity_edges(merge_equivalent_boundary_nodes)
                self._NOISE = _noise.CodeCapacity(self._EDGES)
            else:
                h = d if window_height is None else window_height
                if noise == 'phenomenological':
                    self._N_EDGES, self._EDGES = self._phenomenological_edges(
                        h, False, merge_equivalent_boundary_nodes=merge_equivalent_boundary_nodes)
                    self._NOISE = _noise.Phenomenological(self._EDGES)
                elif noise == 'circuit-level':
                    self._N_EDGES, self._EDGES, families, merges = self._circuit_level_edges(
                        h=h, future_boundary=False,
                        _merge_redundant_edges=_merge_redundant_edges,
                        merge_equivalent_boundary_nodes=merge_equivalent_boundary_nodes)
                    self._NOISE = _noise.CircuitLevel(
                        fresh_edge_dict=families, fresh_merges=merges, **cl_kw)
                else:
                    raise ValueError(f"Unknown noise model {noise!r}.")
            self._SCHEME = (_schemes.Batch if scheme == 'batch' else _schemes.Global)(self, h)
        elif scheme in ('forward', 'frugal'):
            if merge_equivalent_boundary_nodes:
                raise NotImplementedError(
                    "Yet to implement merging equivalent boundary nodes for the stream decoding schemes.")
            if window_height is not None:
                raise ValueError(f"Cannot specify `window_height` for the {scheme} decoding scheme.")
            if scheme == 'forward':
                c = d if commit_height is None else commit_height
                b = d if buffer_height is None else buffer_height
                scheme_class = _schemes.Forward
            else:
                c = 1 if commit_height is None else commit_height
                b = 2 * (d // 2) if buffer_height is None else buffer_height
                scheme_class = _schemes.Frugal
            h = c + b
            if noise == 'code capacity':
                raise TypeError(f"Code capacity incompatible with the {scheme} decoding scheme.")
            elif noise == 'phenomenological':
                self._N_EDGES, self._EDGES = self._phenomenological_edges(h, True)
                _, commit_edges = self._phenomenological_edges(c, True)
                _, fresh = self._phenomenological_edges(c, True, t_start=b)
                self._NOISE = _noise.Phenomenological(fresh, all_edges=self._EDGES)
            elif noise == 'circuit-level':
                self._N_EDGES, self._EDGES, all_families, all_merges = self._circuit_level_edges(
                    h=h, future_boundary=True, _merge_redundant_edges=_merge_redundant_edges)
                _, commit_edges, *_ = self._circuit_level_edges(
                    h=c, future_boundary=True, _merge_redundant_edges=_merge_redundant_edges)
                *_, fresh_families, fresh_merges = self._circuit_level_edges(
                    h=c, future_boundary=True, _merge_redundant_edges=_merge_redundant_edges, t_start=b)
                self._NOISE = _noise.CircuitLevel(
                    fresh_edge_dict=fresh_families, fresh_merges=fresh_merges,
                    all_edge_dict=all_families, all_merges=all_merges, **cl_kw)
            else:
                raise ValueError(f"Unknown noise model {noise!r}.")
            self._SCHEME = scheme_class(self, c, b, commit_edges)
        else:
            raise ValueError(f"Unknown decoding scheme {scheme!r}.")

        incident: dict[Node, set[Edge]] = {}
        for e in self._EDGES:
            for v in e:
                incident.setdefault(v, set()).add(e)
        self._INCIDENT_EDGES = incident
        self._NODES = self._order_nodes(incident.keys())

    # ------------------------------------------------------------------ edges
    @abc.abstractmethod
    def _code_capacity_edges(self, merge_equivalent_boundary_nodes: bool): ...

    @abc.abstractmethod
    def _phenomenological_edges(self, h, future_boundary, t_start=0,
                                merge_equivalent_boundary_nodes=False): ...

    @abc.abstractmethod
    def _circuit_level_edges(self, h, future_boundary, _merge_redundant_edges,
                             t_start=0, merge_equivalent_boundary_nodes=False): ...

    @abc.abstractmethod
    def _future_boundary_nodes(self, h: int) -> list[Node]: ...

    @abc.abstractmethod
    def _redundant_boundary_nodes(self, h: int) -> list[Node]: ...

    @abc.abstractmethod
    def index_to_id(self, index: Node) -> int:
        """Unique node id; all boundary ids lie below all detector ids."""

    def _order_nodes(self, nodes) -> tuple[Node, ...]:
        """Nodes ranked by (boundary first, ``index_to_id``).  For every graph
        but the 1-D repetition code (whose east boundary has the highest id,
        ``codes.py:80-82``) the rank IS the reference id."""
        nodes = list(nodes)
        order = sorted(nodes, key=lambda v: (not self.is_boundary(v), self.index_to_id(v), v))
        self._ID_IS_DENSE = all(self.index_to_id(v) == k for k, v in enumerate(order))
        return tuple(order)

    # ------------------------------------------------------------- properties
    def __repr__(self) -> str:
        return f"{type(self).__name__}({self.D}, '{self.NOISE}', scheme='{self.SCHEME}')"

    @property
    def D(self): return self._D
    @property
    def MERGED_EQUIVALENT_BOUNDARY_NODES(self): return self._MERGED_EQUIVALENT_BOUNDARY_NODES
    @property
    def SCHEME(self): return self._SCHEME
    @property
    def N_EDGES(self): return self._N_EDGES
    @property
    def EDGES(self): return self._EDGES
    @property
    def NODES(self): return self._NODES
    @property
    def NOISE(self): return self._NOISE
    @property
    def TIME_AXIS(self): return self._TIME_AXIS
    @property
    def LONG_AXIS(self): return self._LONG_AXIS
    @property
    def INCIDENT_EDGES(self): return self._INCIDENT_EDGES

    @cached_property
    def NODE_COUNT(self): return len(self.NODES)

    @cached_property
    def DIMENSION(self):
        return self._CODE_DIMENSION + (str(self.NOISE) != 'code capacity')

    @cached_property
    def DETECTORS(self):
        return tuple(v for v in self.NODES if not self.is_boundary(v))

    @cached_property
    def DETECTOR_COUNT(self): return len(self.DETECTORS)

    @cached_property
    def BOUNDARY_NODES(self):
        return tuple(v for v in self.NODES if self.is_boundary(v))

    @cached_property
    def FLAT(self) -> FlatGraph:
        """Integer lowering of this graph for the device (built once)."""
        return FlatGraph.from_code(self)

    # ---------------------------------------------------------------- helpers
    def is_boundary(self, v: Node) -> bool:
        return self.SCHEME.is_boundary(v)

    def neighbors(self, v: Node) -> set[Node]:
        out = {w for e in self.INCIDENT_EDGES[v] for w in e}
        out.discard(v)
        return out

    @staticmethod
    def traverse_edge(e: Edge, u: Node) -> Node:
        return e[1] if u == e[0] else e[0]

    def raise_node(self, v: Node, delta_t: int = 1) -> Node:
        return _shift(v, self.TIME_AXIS, delta_t)

    def raise_edge(self, e: Edge, delta_t: int = 1) -> Edge:
        return (self.raise_node(e[0], delta_t), self.raise_node(e[1], delta_t))

    def make_error(self, noise_level: float, exclude_future_boundary: bool = False) -> set[Edge]:
        """Host sampler (``_base_classes.py:400-425``): consumes Python's global
        ``random`` stream exactly like the reference, so ``random.seed(s)``
        reproduces the reference's error sets.  The Monte-Carlo hot path does
        NOT come through here -- it samples on the device (kernel K1)."""
        error = self.NOISE.make_error(noise_level)
        if exclude_future_boundary:
            top = self.SCHEME.WINDOW_HEIGHT
            error = {e for e in error if all(v[self.TIME_AXIS] != top for v in e)}
        return error

    def get_verbose_syndrome(self, error) -> set[Node]:
        out: set[Node] = set()
        for e in error:
            out.symmetric_difference_update(e)
        return out

    def get_syndrome(self, error) -> set[Node]:
        return {v for v in self.get_verbose_syndrome(error) if not self.is_boundary(v)}

    @staticmethod
    def compose_errors(*errors) -> set[Edge]:
        out: set[Edge] = set()
        for e in errors:
            out ^= e
        return out

    def get_logical_error(self, leftover) -> int:
        return self.SCHEME.get_logical_error(leftover)


# ---------------------------------------------------------------------------
def _grid(*ranges) -> Iterator[tuple]:
    """Row-major product of ranges (last index fastest)."""
    if not ranges:
        yield ()
        return
    head, *tail = ranges
    for x in head:
        for rest in _grid(*tail):
            yield (x, *rest)


def _family(delta, *ranges) -> tuple[Edge, ...]:
    """All edges ``(v, v+delta)`` for ``v`` in the row-major grid of ``ranges``."""
    return tuple((v, tuple(a + b for a, b in zip(v, delta))) for v in _grid(*ranges))


class Repetition(Code):
    """Decoding graph of the distance-``d`` repetition code (``codes.py:19-95``)."""

    _LONG_AXIS = 0
    _CODE_DIMENSION = 1

    def _code_capacity_edges(self, merge_equivalent_boundary_nodes):
        d = self.D
        return d, _family((1,), range(-1, d - 1))

    def _phenomenological_edges(self, h, future_boundary, t_start=0,
                                merge_equivalent_boundary_nodes=False):
        d = self.D
        layers = h if future_boundary else h - 1
        ts = range(t_start, t_start + h)
        if merge_equivalent_boundary_nodes:
            horizontal = (
                tuple(((-1, 0), (0, t)) for t in ts)
                + _family((1, 0), range(d - 2), ts)
                + tuple(((d - 2, t), (d - 1, 0)) for t in ts)
            )
        else:
            horizontal = _family((1, 0), range(-1, d - 1), ts)
        vertical = _family((0, 1), range(d - 1), range(t_start, t_start + layers))
        return h * d + layers * (d - 1), horizontal + vertical

    def _future_boundary_nodes(self, h):
        return [(j, h) for j in range(self.D - 1)]

    def _redundant_boundary_nodes(self, h):
        raise NotImplementedError("Yet to implement circuit-level noise for repetition code.")

    def _circuit_level_edges(self, **_):
        raise NotImplementedError("Yet to implement circuit-level noise for repetition code.")

    def index_to_id(self, index):
        d = self.D
        if len(index) == 1:
            return index[0] + 1
        h = self.SCHEME.WINDOW_HEIGHT
        j, t = index
        if j == -1:
            return 2 * t
        if j == d - 1:
            return 2 * t + 1
        if t == h:
            return 2 * h + j
        base = 2 * h if isinstance(self.SCHEME, _schemes.Batch) else 2 * h + d - 1
        return base + (d - 1) * t + j


class Surface(Code):
    """Decoding graph of the unrotated surface code (``codes.py:98-324``)."""

    _LONG_AXIS = 1
    _CODE_DIMENSION = 2

    @property
    def DATA_QUBIT_COUNT(self) -> int:
        d = self.D
        return d * d + (d - 1) * (d - 1)

    def _code_capacity_edges(self, merge_equivalent_boundary_nodes):
        d = self.D
        south = _family((1, 0), range(d - 1), range(d - 1))
        if merge_equivalent_boundary_nodes:
            east = (
                tuple(((0, -1), (i, 0)) for i in range(d))
                + _family((0, 1), range(d), range(d - 2))
                + tuple(((i, d - 2), (0, d - 1)) for i in range(d))
            )
        else:
            east = _family((0, 1), range(d), range(-1, d - 1))
        return self.DATA_QUBIT_COUNT, south + east

    def _phenomenological_edges(self, h, future_boundary, t_start=0,
                                merge_equivalent_boundary_nodes=False):
        d = self.D
        layers = h if future_boundary else h - 1
        ts = range(t_start, t_start + h)
        south = _family((1, 0, 0), range(d - 1), range(d - 1), ts)
        if merge_equivalent_boundary_nodes:
            east = (
                tuple(((0, -1, 0), (i, 0, t)) for i in range(d) for t in ts)
                + _family((0, 1, 0), range(d), range(d - 2), ts)
                + tuple(((i, d - 2, t), (0, d - 1, 0)) for i in range(d) for t in ts)
            )
        else:
            east = _family((0, 1, 0), range(d), range(-1, d - 1), ts)
        up = _family((0, 0, 1), range(d), range(d - 1), range(t_start, t_start + layers))
        return h * self.DATA_QUBIT_COUNT + layers * d * (d - 1), south + east + up

    def _future_boundary_nodes(self, h):
        d = self.D
        return [(i, j, h) for i in range(d) for j in range(d - 1)]

    def _redundant_boundary_nodes(self, h):
        d = self.D
        return [(i, j, t) for i in range(d) for j, t in ((-1, -1), (d - 1, h))]

    def _circuit_level_edges(self, h, future_boundary, _merge_redundant_edges,
                             t_start=0, merge_equivalent_boundary_nodes=False):
        """Twelve edge families of the circuit-level graph (``codes.py:155-251``).

        Columns are split west-wall / bulk / east-wall because diagonal edges
        that touch a space boundary are redundant with the horizontal boundary
        edge of the detector's layer; when merging is on they are dropped from
        ``EDGES`` and only add their fault multiplicity to the substitute.
        """
        d = self.D
        t0 = t_start
        west, bulk, east = (-1,), range(d - 2), (d - 2,)
        if future_boundary:
            layers = h
            diag_t = (range(t0 - 1, t0 - 1 + layers), range(t0, t0 + layers), range(t0, t0 + layers))
        else:
            layers = h - 1
            diag_t = (range(t0, t0 + layers),) * 3
        cols = tuple(zip((west, bulk, east), diag_t))
        ts = range(t0, t0 + h)
        rim, core = (0, d - 1), range(1, d - 1)

        s = _family((1, 0, 0), range(d - 1), range(d - 1), ts)
        e_wm, e_bulk, e_em = (_family((0, 1, 0), range(d), js, ts) for js in (west, bulk, east))
        if merge_equivalent_boundary_nodes:
            e_wm = tuple(((0, -1, 0), (i, 0, t)) for i in range(d) for t in ts)
            e_em = tuple(((i, d - 2, t), (0, d - 1, 0)) for i in range(d) for t in ts)
        u3, u4 = (_family((0, 0, 1), rows, range(d - 1), range(t0, t0 + layers)) for rows in (rim, core))
        sd = _family((1, 0, -1), range(d - 1), range(d - 1), range(t0 + 1, t0 + 1 + layers))
        eu_wc, eu_ns, eu_ec = (_family((0, 1, 1), rim, js, tr) for js, tr in cols)
        eu_w, eu_centre, eu_e = (_family((0, 1, 1), core, js, tr) for js, tr in cols)
        eu_edge = eu_ns + eu_w + eu_e
        seu_w, seu_bulk, seu_e = (_family((1, 1, 1), range(d - 1), js, tr) for js, tr in cols)
        seu = seu_w + seu_e + seu_bulk

        if _merge_redundant_edges or merge_equivalent_boundary_nodes:
            n_edges = h * self.DATA_QUBIT_COUNT + layers * (2 * d - 1) * (2 * d - 3)
            edges = s + e_wm + e_bulk + e_em + u3 + u4 + sd + eu_ns + eu_centre + seu_bulk
            merges = {e: self._substitute(e, merge_equivalent_boundary_nodes)
                      for e in eu_wc + eu_ec + eu_w + eu_e + seu_w + seu_e}
        else:
            n_edges = (h + layers) * self.DATA_QUBIT_COUNT + 2 * d * (d - 1) * layers
            edges = s + e_wm + e_bulk + e_em + u3 + u4 + sd + eu_wc + eu_ec + eu_edge + eu_centre + seu
            merges = None
        families = {
            'S': s, 'E westmost': e_wm, 'E bulk': e_bulk, 'E eastmost': e_em,
            'U 3': u3, 'U 4': u4, 'SD': sd,
            'EU west corners': eu_wc, 'EU east corners': eu_ec, 'EU edge': eu_edge,
            'EU centre': eu_centre, 'SEU': seu,
        }
        return n_edges, edges, families, merges

    def _substitute(self, e: Edge, merge_equivalent_boundary_nodes: bool = False) -> Edge:
        """Horizontal boundary edge standing in for the redundant diagonal ``e``
        (``codes.py:253-268``)."""
        d, a = self.D, self.LONG_AXIS
        u, v = e

        def wall(coord: int, like: Node) -> Node:
            if merge_equivalent_boundary_nodes:
                return tuple(coord if k == a else 0 for k in range(len(like)))
            return tuple(coord if k == a else x for k, x in enumerate(like))

        if u[a] == -1:
            return (wall(-1, v), v)
        if v[a] == d - 1:
            return (u, wall(d - 1, u))
        raise ValueError(f'Edge {e} must have a boundary node.')

    def index_to_label(self, index: Node):
        if self.DIMENSION != 2:
            raise NotImplementedError('Only implemented for code capacity.')
        i, j = index
        return (self.D + 1) * i + j + 1

    def label_to_index(self, a: int) -> Node:
        if self.DIMENSION != 2:
            raise NotImplementedError('Only implemented for code capacity.')
        return (a // (self.D + 1), a % (self.D + 1) - 1)

    def index_to_id(self, index):
        d = self.D
        if len(index) == 2:
            i, j = index
            if j == -1:
                return 2 * i
            if j == d - 1:
                return 2 * i + 1
            return 2 * d + (d - 1) * i + j
        h = self.SCHEME.WINDOW_HEIGHT
        i, j, t = index
        if j == -1:
            return 2 * h * i + t
        if j == d - 1:
            return (2 * i + 1) * h + t
        if t == h:
            return 2 * d * h + (d - 1) * i + j
        base = 2 * d * h if isinstance(self.SCHEME, _schemes.Batch) else 2 * d * h + d * (d - 1)
        return base + h * (d - 1) * i + h * j + t
