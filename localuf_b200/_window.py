"""Layered (per time-layer) tables of a streaming window, for Snowflake and the
frugal scheme.

The reference addresses the sliding window through coordinate tuples and
rebuilds Python sets every cycle (``_schemes.py:532-564``,
``decoders/snowflake/main.py:298-322,1161-1211,1547-1564``).  Here a window of
``h`` layers is two arrays of identical *layer blocks*:

* nodes: in-layer position ``q`` (``q < NLb`` space-boundary nodes, then the
  detectors) and layer ``t``; Snowflake's own time-reversed id
  (``snowflake/main.py:159-179``) is ``NLb*(h-1-t) + q`` for boundary nodes and
  ``NLb*h + NLd*(h-1-t) + (q-NLb)`` for detectors, so that "drop one layer" is
  ``id += NLb`` / ``id += NLd`` (``id_below``, ``:186-199``);
* edges: block ``tmin`` (lowest layer touched; block 0 is the floor sheet that
  is committed at the next drop, block ``h-1`` is the freshly sampled sheet,
  whose upward edges end on the future boundary ``t = h``) and a local index
  ``l``; layered index ``(h-1-tmin)*EL + l``.  ``l`` is the rank of the edge
  among the floor edges *in code-edge-id order*, so scanning a block in index
  order is the canonical order in which floor edges are handed to ``Pairs``.

Every table below is per in-layer position / per local edge: the lattice is
translation invariant in time.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np

_REP_ORDER = ('W', 'E', 'D', 'U')
_SURF_ORDER = ('NWD', 'N', 'NU', 'WD', 'W', 'D', 'U', 'E', 'EU', 'SD', 'S', 'SEU')
_REP_DELTA = {'W': (-1, 0), 'E': (1, 0), 'D': (0, -1), 'U': (0, 1)}
_SURF_DELTA = {'NWD': (-1, -1, -1), 'N': (-1, 0, 0), 'NU': (-1, 0, 1), 'WD': (0, -1, -1), 'W': (0, -1, 0),
               'D': (0, 0, -1), 'U': (0, 0, 1), 'E': (0, 1, 0), 'EU': (0, 1, 1), 'SD': (1, 0, -1),
               'S': (1, 0, 0), 'SEU': (1, 1, 1)}


@dataclass
class WindowTables:
    d: int
    h: int
    NLb: int                     # space-boundary nodes per layer
    NLd: int                     # detectors per layer
    EL: int                      # edges per layer block
    K: int                       # neighbour slots
    dir_names: tuple
    nb_valid: np.ndarray         # uint8[NL, K]
    nb_q: np.ndarray             # int32[NL, K]  neighbour's in-layer position
    nb_dt: np.ndarray            # int32[NL, K]  neighbour's layer - node's layer
    nb_el: np.ndarray            # int32[NL, K]  local index of the connecting edge
    nb_et: np.ndarray            # int32[NL, K]  edge block - node's layer (0 or -1)
    edge_q: np.ndarray           # int32[EL, 2]  endpoint positions (first, second endpoint)
    edge_dt: np.ndarray          # int32[EL, 2]  endpoint layer - edge block (0 or 1)
    edge_class: np.ndarray       # uint8[EL]     probability class of the edge when freshly sampled
    edge_fb: np.ndarray          # uint8[EL]     touches the future boundary when in the top block
    pos_long: np.ndarray         # int32[NL]     coordinate along the long axis
    code_edge: np.ndarray        # int32[h*EL]   layered index -> position in code.EDGES
    layered_of_code: np.ndarray  # int32[E]      position in code.EDGES -> layered index
    n_classes: int
    rep: bool = False            # Repetition (one space axis) or Surface (two)

    @property
    def NL(self):
        return self.NLb + self.NLd

    def position_of(self, v) -> int:
        """In-layer position of node ``v`` (its time coordinate is ignored)."""
        d = self.d
        if self.rep:
            j, = v[:-1]
            return 0 if j == -1 else 1 if j == d - 1 else 2 + j
        i, j = v[:-1]
        return i if j == -1 else d + i if j == d - 1 else 2 * d + d * j + i

    def node_id(self, q: int, t: int) -> int:
        if q < self.NLb:
            return self.NLb * (self.h - 1 - t) + q
        return self.NLb * self.h + self.NLd * (self.h - 1 - t) + (q - self.NLb)


def build(code, neighbor_order=None) -> WindowTables:
    """Tables of a frugal-scheme code (commit height 1)."""
    scheme = code.SCHEME
    if getattr(scheme, '_COMMIT_HEIGHT', None) != 1:
        raise NotImplementedError("The streaming engine assumes a commit height of 1 (the frugal default).")
    d, h, ta, a = code.D, scheme.WINDOW_HEIGHT, code.TIME_AXIS, code.LONG_AXIS
    rep = code._CODE_DIMENSION == 1
    order = tuple(neighbor_order) if neighbor_order is not None else (_REP_ORDER if rep else _SURF_ORDER)
    delta = _REP_DELTA if rep else _SURF_DELTA

    def pos_of(v):               # in-layer position of a node (time coordinate dropped)
        if rep:
            j, = v[:-1]
            return 0 if j == -1 else 1 if j == d - 1 else 2 + j
        i, j = v[:-1]
        return i if j == -1 else d + i if j == d - 1 else 2 * d + d * j + i

    NLb, NLd = (2, d - 1) if rep else (2 * d, d * (d - 1))
    NL = NLb + NLd
    space = [None] * NL           # position -> spatial coordinates
    for v in code.NODES:
        if v[ta] == 0:
            space[pos_of(v)] = v[:-1]
    assert all(s is not None for s in space)

    edges = code.EDGES
    eidx = {e: k for k, e in enumerate(edges)}
    tmin = [min(e[0][ta], e[1][ta]) for e in edges]
    floor = [k for k in range(len(edges)) if tmin[k] == 0]           # ascending code edge id
    EL = len(floor)
    if EL * h != len(edges):
        raise AssertionError("window is not a stack of identical layer blocks")
    local_of_pattern = {}
    edge_q = np.zeros((EL, 2), np.int32)
    edge_dt = np.zeros((EL, 2), np.int32)
    for l, k in enumerate(floor):
        u, v = edges[k]
        local_of_pattern[(pos_of(u), u[ta], pos_of(v), v[ta])] = l
        edge_q[l] = (pos_of(u), pos_of(v))
        edge_dt[l] = (u[ta], v[ta])
    layered_of_code = np.zeros(len(edges), np.int32)
    code_edge = np.zeros(h * EL, np.int32)
    for k, (u, v) in enumerate(edges):
        l = local_of_pattern[(pos_of(u), u[ta] - tmin[k], pos_of(v), v[ta] - tmin[k])]
        idx = (h - 1 - tmin[k]) * EL + l
        layered_of_code[k] = idx
        code_edge[idx] = k
    assert sorted(layered_of_code.tolist()) == list(range(h * EL))

    # probability class / future-boundary flag of the freshly sampled block (tmin = h-1)
    edge_class = np.full(EL, 255, np.uint8)
    edge_fb = np.zeros(EL, np.uint8)
    classes = code.NOISE.classes()
    for c, members in enumerate(classes):
        for e in members:
            k = eidx[e]
            assert tmin[k] == h - 1, "fresh edges must be the top layer block"
            l = layered_of_code[k] - 0 * EL
            assert layered_of_code[k] < EL
            edge_class[l] = c
            edge_fb[l] = 1 if (e[0][ta] == h or e[1][ta] == h) else 0
    assert not np.any(edge_class == 255)

    K = len(order)
    nb_valid = np.zeros((NL, K), np.uint8)
    nb_q = np.zeros((NL, K), np.int32)
    nb_dt = np.zeros((NL, K), np.int32)
    nb_el = np.zeros((NL, K), np.int32)
    nb_et = np.zeros((NL, K), np.int32)
    t_probe = 1 if h >= 3 else 0      # a layer with both a layer below and a layer above
    for q in range(NL):
        v = (*space[q], t_probe)
        for s, name in enumerate(order):
            dl = delta[name]
            w = tuple(x + y for x, y in zip(v, dl))
            forward = dl > tuple(0 for _ in dl)
            e = (v, w) if forward else (w, v)
            if e not in eidx:
                continue
            k = eidx[e]
            nb_valid[q, s] = 1
            nb_q[q, s] = pos_of(w)
            nb_dt[q, s] = dl[-1]
            nb_el[q, s] = layered_of_code[k] % EL
            nb_et[q, s] = tmin[k] - t_probe
    if h < 3:
        raise NotImplementedError("window height below 3 is not supported by the streaming engine")

    pos_long = np.array([(-1 if (q < NLb and (q == 0 if rep else q < d)) else d - 1) if q < NLb else space[q][a if not rep else 0]
                         for q in range(NL)], np.int32)
    return WindowTables(d=d, h=h, NLb=NLb, NLd=NLd, EL=EL, K=K, dir_names=order,
                        nb_valid=nb_valid, nb_q=nb_q, nb_dt=nb_dt, nb_el=nb_el, nb_et=nb_et,
                        edge_q=edge_q, edge_dt=edge_dt, edge_class=edge_class, edge_fb=edge_fb,
                        pos_long=pos_long, code_edge=code_edge, layered_of_code=layered_of_code,
                        n_classes=len(classes), rep=rep)


# ---------------------------------------------------------------------------
@dataclass
class ForwardTables:
    """Window bookkeeping of the forward scheme (``_schemes.py:282-450``) over
    the flat graph of the same code: edge / node indices are those of
    :class:`localuf_b200._graph.FlatGraph`."""
    c: int                       # commit height
    b: int                       # buffer height
    n_cleanse: int               # h // c noiseless cycles at the end
    lower_edge: np.ndarray       # int32[E]   edge lowered by c layers (-1 if it leaves the window)
    commit_mask: np.ndarray      # uint32[W(E)]  edges of the commit region
    art_node: np.ndarray         # int32[E, 2]  per endpoint: the detector it makes artificial after the commit
                                 #              (endpoint on layer t == c, lowered to t == 0), else -1
    edge_fb: np.ndarray          # uint8[E]   edge touches the future boundary t == h
    node_lower: np.ndarray       # int32[N]   node lowered by c layers (-1 if it leaves the window)
    node_side: np.ndarray        # uint8[N]   1 west boundary, 2 east boundary, 0 otherwise
    buffer_class: np.ndarray     # uint8[E]   probability class of a buffer-region edge (its fresh translate's), 255 in the commit region


def build_forward(code) -> ForwardTables:
    scheme = code.SCHEME
    flat = code.FLAT
    c, b = scheme._COMMIT_HEIGHT, scheme._BUFFER_HEIGHT
    h, ta, a, d = c + b, code.TIME_AXIS, code.LONG_AXIS, code.D
    E, N = flat.n_edges, flat.n_nodes
    lower_edge = np.full(E, -1, np.int32)
    art_node = np.full((E, 2), -1, np.int32)
    edge_fb = np.zeros(E, np.uint8)
    commit = set(scheme._COMMIT_EDGES)
    commit_idx = []
    for k, e in enumerate(flat.edges):
        low = code.raise_edge(e, -c)
        lower_edge[k] = flat.edge_index.get(low, -1)
        edge_fb[k] = 1 if any(v[ta] == h for v in e) else 0
        if e in commit:
            commit_idx.append(k)
            for x, v in enumerate(e):
                if v[ta] == c and not code.is_boundary(v):
                    art_node[k, x] = flat.node_index[code.raise_node(v, -c)]
    from localuf_b200._graph import pack_bits
    buffer_class = np.full(E, 255, np.uint8)
    for k, e in enumerate(flat.edges):
        if e in commit:
            continue
        up = e
        while flat.edge_class[flat.edge_index[up]] == 255:      # translate upwards into the fresh sheet
            up = code.raise_edge(up, c)
        buffer_class[k] = flat.edge_class[flat.edge_index[up]]
    node_lower = np.full(N, -1, np.int32)
    node_side = np.zeros(N, np.uint8)
    for k, v in enumerate(flat.nodes):
        node_lower[k] = flat.node_index.get(code.raise_node(v, -c), -1)
        node_side[k] = 1 if v[a] == -1 else 2 if v[a] == d - 1 else 0
    return ForwardTables(c=c, b=b, n_cleanse=h // c, lower_edge=lower_edge,
                         commit_mask=pack_bits(commit_idx, E), art_node=art_node, edge_fb=edge_fb,
                         node_lower=node_lower, node_side=node_side, buffer_class=buffer_class)
