"""Flat (integer-array) lowering of a decoding graph for the device.

The reference never leaves Python sets/dicts keyed by coordinate tuples
(SURVEY.md section 1).  Everything the CUDA kernels need is derived here once
per code object and uploaded once per ``luf_graph``:

* node index = position in ``code.NODES``: all boundary nodes first
  (``[0, B)``), detectors after (``[B, N)``); equal to the reference's
  ``index_to_id`` (``codes.py:64-82,287-311``) whenever that is dense;
* edge index = position in ``code.EDGES`` (``_base_classes.py:165``);
  ``edge_u[e]`` is the endpoint stored first by the reference (north / west /
  down side), which is what the logical check looks at
  (``_schemes.py:123-129``) and what breaks union-by-size ties
  (``decoders/uf.py:270-273``);
* CSR incidence sorted by edge id (the canonical neighbour order of the UF
  peeling traversal);
* a direction-ordered neighbour table for the cellular-automaton decoders
  (``decoders/luf/main.py:743-783``);
* bit-packed masks (32-bit words, little-endian bit order: bit ``k`` of word
  ``w`` is element ``32*w + k``).
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

NODE_BOUNDARY = 1   # any boundary (space or future)
NODE_WEST = 2       # LONG_AXIS coordinate == -1
NODE_EAST = 4       # LONG_AXIS coordinate == d-1
NODE_FUTURE = 8     # TIME_AXIS coordinate == window height (streaming schemes)

NOT_FRESH = 255

# neighbour scan orders of the CA decoders, decoders/luf/main.py:748-777
_DIRS_2D = (('N', (-1, 0)), ('W', (0, -1)), ('E', (0, 1)), ('S', (1, 0)))
_DIRS_PHEN = (('N', (-1, 0, 0)), ('W', (0, -1, 0)), ('D', (0, 0, -1)),
              ('U', (0, 0, 1)), ('E', (0, 1, 0)), ('S', (1, 0, 0)))
_DIRS_CL = (('NWD', (-1, -1, -1)), ('N', (-1, 0, 0)), ('NU', (-1, 0, 1)), ('WD', (0, -1, -1)),
            ('W', (0, -1, 0)), ('D', (0, 0, -1)), ('U', (0, 0, 1)), ('E', (0, 1, 0)),
            ('EU', (0, 1, 1)), ('SD', (1, 0, -1)), ('S', (1, 0, 0)), ('SEU', (1, 1, 1)))


def words(nbits: int) -> int:
    """Number of 32-bit words holding ``nbits`` bits."""
    return (nbits + 31) // 32


def pack_bits(indices, nbits: int) -> np.ndarray:
    out = np.zeros(words(nbits), dtype=np.uint32)
    for k in indices:
        out[k >> 5] |= np.uint32(1 << (k & 31))
    return out


def unpack_bits(word_array, nbits: int) -> list[int]:
    w = np.ascontiguousarray(word_array, dtype=np.uint32)
    bits = np.unpackbits(w.view(np.uint8), bitorder='little')[:nbits]
    return np.flatnonzero(bits).tolist()


@dataclass
class FlatGraph:
    n_nodes: int
    n_boundary: int
    n_edges: int
    d: int
    height: int
    edge_u: np.ndarray            # int32[E]
    edge_v: np.ndarray            # int32[E]
    node_flags: np.ndarray        # uint8[N]
    node_id: np.ndarray           # int32[N]  reference index_to_id (initial cluster id of the CA decoders)
    node_long: np.ndarray         # int32[N]  coordinate along LONG_AXIS (-1 west ... d-1 east)
    inc_off: np.ndarray           # int32[N+1]
    inc_edge: np.ndarray          # int32[2E]  ascending edge id within a node
    inc_other: np.ndarray         # int32[2E]
    west_edge_mask: np.ndarray    # uint32[W(E)]  edges whose first endpoint is on the west boundary
    edge_class: np.ndarray        # uint8[E]  probability class, NOT_FRESH if never sampled
    n_classes: int
    dir_names: tuple              # direction labels of the CA neighbour table
    nbr_edge: np.ndarray          # int32[N, K]  edge id in scan order, -1 if absent
    nbr_node: np.ndarray          # int32[N, K]  neighbour node index, -1 if absent
    nbr_owned: np.ndarray         # uint8[N, K]  1 if this node is the edge's first endpoint
    actis_span: np.ndarray        # int32[N]  per-node countdown start of Actis (luf/main.py:970-976)
    actis_relay: np.ndarray       # int32[N]  neighbour towards node id 0 (luf/main.py:1138-1163), -1 = the controller
    actis_global_span: int        # controller countdown start (luf/main.py:517-523)
    dense_ids: bool               # node index == reference index_to_id
    node_index: dict = field(repr=False, default_factory=dict)   # coord -> node index
    edge_index: dict = field(repr=False, default_factory=dict)   # edge tuple -> edge index
    nodes: tuple = field(repr=False, default=())
    edges: tuple = field(repr=False, default=())

    @property
    def n_detectors(self) -> int:
        return self.n_nodes - self.n_boundary

    @property
    def max_degree(self) -> int:
        return int(np.max(np.diff(self.inc_off))) if self.n_nodes else 0

    # ------------------------------------------------------------------
    @classmethod
    def from_code(cls, code) -> 'FlatGraph':
        nodes = code.NODES
        edges = code.EDGES
        N, E = len(nodes), len(edges)
        node_index = {v: k for k, v in enumerate(nodes)}
        edge_index = {e: k for k, e in enumerate(edges)}
        a, ta, d = code.LONG_AXIS, code.TIME_AXIS, code.D
        h = code.SCHEME.WINDOW_HEIGHT
        streaming = hasattr(code.SCHEME, '_COMMIT_HEIGHT')

        flags = np.zeros(N, dtype=np.uint8)
        for k, v in enumerate(nodes):
            f = 0
            if code.is_boundary(v):
                f |= NODE_BOUNDARY
            if v[a] == -1:
                f |= NODE_WEST
            elif v[a] == d - 1:
                f |= NODE_EAST
            if streaming and code.DIMENSION > code._CODE_DIMENSION and v[ta] == h:
                f |= NODE_FUTURE
            flags[k] = f
        n_boundary = int(np.count_nonzero(flags & NODE_BOUNDARY))
        if not np.all(flags[:n_boundary] & NODE_BOUNDARY):
            raise AssertionError("boundary nodes must occupy the lowest node indices")

        eu = np.fromiter((node_index[e[0]] for e in edges), dtype=np.int32, count=E)
        ev = np.fromiter((node_index[e[1]] for e in edges), dtype=np.int32, count=E)

        deg = np.bincount(eu, minlength=N) + np.bincount(ev, minlength=N)
        off = np.zeros(N + 1, dtype=np.int32)
        np.cumsum(deg, out=off[1:])
        cursor = off[:-1].copy()
        inc_edge = np.empty(2 * E, dtype=np.int32)
        inc_other = np.empty(2 * E, dtype=np.int32)
        for k in range(E):              # ascending edge id => each node's slice is sorted
            for me, other in ((eu[k], ev[k]), (ev[k], eu[k])):
                inc_edge[cursor[me]] = k
                inc_other[cursor[me]] = other
                cursor[me] += 1

        west = pack_bits((k for k in range(E) if flags[eu[k]] & NODE_WEST), E)

        edge_class = np.full(E, NOT_FRESH, dtype=np.uint8)
        class_lists = code.NOISE.classes()
        if len(class_lists) >= NOT_FRESH:
            raise ValueError("too many probability classes")
        for c, members in enumerate(class_lists):
            for e in members:
                edge_class[edge_index[e]] = c

        dims = len(nodes[0]) if N else 0
        if dims == 2 and str(code.NOISE) == 'code capacity' and code._CODE_DIMENSION == 2:
            dirs = _DIRS_2D
        elif dims == 3 and str(code.NOISE) == 'phenomenological':
            dirs = _DIRS_PHEN
        elif dims == 3:
            dirs = _DIRS_CL
        elif dims == 1:     # repetition, code capacity: W, E
            dirs = (('W', (-1,)), ('E', (1,)))
        else:               # repetition, phenomenological: W, E, D, U (snowflake/main.py:86-91)
            dirs = (('W', (-1, 0)), ('E', (1, 0)), ('D', (0, -1)), ('U', (0, 1)))
        K = len(dirs)
        nbr_edge = np.full((N, K), -1, dtype=np.int32)
        nbr_node = np.full((N, K), -1, dtype=np.int32)
        nbr_owned = np.zeros((N, K), dtype=np.uint8)
        for k, v in enumerate(nodes):
            for s, (_, delta) in enumerate(dirs):
                w = tuple(x + y for x, y in zip(v, delta))
                forward = delta > tuple(0 for _ in delta)   # first non-zero component positive
                e = (v, w) if forward else (w, v)
                if e in edge_index and w in node_index:
                    nbr_edge[k, s] = edge_index[e]
                    nbr_node[k, s] = node_index[w]
                    nbr_owned[k, s] = 1 if forward else 0

        span = np.zeros(N, dtype=np.int32)
        relay = np.full(N, -1, dtype=np.int32)
        global_span = 0
        kind = str(code.NOISE)
        if code._CODE_DIMENSION == 2 and not streaming and bool(getattr(code, '_ID_IS_DENSE', False)):
            global_span = {'code capacity': 2 * d, 'phenomenological': 3 * d - 1}.get(kind, d + 1)
            for k, v in enumerate(nodes):
                if kind == 'circuit-level':
                    i, j, t = v
                    span[k] = d - max(i, j + 1, t)
                    w = (i - 1 if i > 0 else i, j - 1 if j > -1 else j, t - 1 if t > 0 else t)
                else:
                    span[k] = code.DIMENSION * (d - 1) - sum(v)
                    if len(v) == 2:
                        i, j = v
                        w = (i, j - 1) if j > -1 else ((i - 1, j) if i > 0 else None)
                    else:
                        i, j, t = v
                        w = (i, j, t - 1) if t > 0 else ((i, j - 1, t) if j > -1 else ((i - 1, j, t) if i > 0 else None))
                relay[k] = -1 if (w is None or w == v) else node_index.get(w, -2)

        return cls(
            n_nodes=N, n_boundary=n_boundary, n_edges=E, d=d, height=h,
            edge_u=eu, edge_v=ev, node_flags=flags,
            node_id=np.fromiter((code.index_to_id(v) for v in nodes), dtype=np.int32, count=N),
            node_long=np.fromiter((v[a] for v in nodes), dtype=np.int32, count=N),
            inc_off=off, inc_edge=inc_edge, inc_other=inc_other,
            west_edge_mask=west, edge_class=edge_class, n_classes=len(class_lists),
            dir_names=tuple(n for n, _ in dirs),
            nbr_edge=nbr_edge, nbr_node=nbr_node, nbr_owned=nbr_owned,
            actis_span=span, actis_relay=relay, actis_global_span=global_span,
            dense_ids=bool(getattr(code, '_ID_IS_DENSE', False)),
            node_index=node_index, edge_index=edge_index, nodes=nodes, edges=edges,
        )

    # ------------------------------------------------------------ conversions
    def edges_to_bits(self, edge_set) -> np.ndarray:
        return pack_bits((self.edge_index[e] for e in edge_set), self.n_edges)

    def bits_to_edges(self, word_array) -> set:
        return {self.edges[k] for k in unpack_bits(word_array, self.n_edges)}

    def syndrome_to_bits(self, node_set) -> np.ndarray:
        """Detector bit ``k`` is node ``B + k``."""
        B = self.n_boundary
        idx = []
        for v in node_set:
            k = self.node_index[v] - B
            if k < 0:
                raise ValueError(f"{v} is a boundary node, not a detector")
            idx.append(k)
        return pack_bits(idx, self.n_detectors)

    def bits_to_syndrome(self, word_array) -> set:
        B = self.n_boundary
        return {self.nodes[B + k] for k in unpack_bits(word_array, self.n_detectors)}
