"""Canonical-order execution of the UNMODIFIED reference UF decoder.

TEST INFRASTRUCTURE ONLY (authoring container; needs ``/root/reference``).

The reference ``UF`` (``localuf/decoders/uf.py``) keeps its active clusters in
a ``set`` hashed by ``id()`` (``uf.py:69-70,217-222``), so the order in which a
round's fully-grown edges are merged -- and through it union tie-breaks
(``:270-273``), which boundary node a cluster keeps (``:536-549``) and the
shape of the peeling tree (``:326-344``) -- changes from process to process.
Growth/erasure do not depend on it; corrections do (hard
part 1).

``CanonicalUF`` subclasses the reference class and pins *only iteration
order*; every rule (``_find``, ``_union``, inclination, forester,
``_update_self_after_union``, ``peel``) is inherited unchanged:

* ``_growth_round``: the round's merge list is processed in ascending edge id
  (position in ``code.EDGES``); growth saturates at FULL instead of raising on
  the (never observed) ``Growth(3)``;
* ``_make_tree``: a node's fully-grown incident edges are scanned in ascending
  edge id.

It is therefore one legal execution of the reference algorithm, and it is the
execution the C oracle and the CUDA kernel reproduce bit for bit.
"""
from __future__ import annotations

import collections

import ref_shim

localuf = ref_shim.load()
from localuf.constants import Growth  # noqa: E402
from localuf.decoders.uf import UF as _RefUF  # noqa: E402


class CanonicalUF(_RefUF):
    def __init__(self, code, dynamic=False, inclination='default'):
        self._EDGE_ID = {e: k for k, e in enumerate(code.EDGES)}
        self.round_count = 0
        super().__init__(code, dynamic=dynamic, inclination=inclination)

    def reset(self, no_boundaries=False):
        super().reset(no_boundaries=no_boundaries)
        self.round_count = 0

    def _growth_round(self, log_history, clusters_to_grow=None):
        if clusters_to_grow is None:
            clusters_to_grow = self.active_clusters
        self.round_count += 1
        merge_ls = []
        for cluster in clusters_to_grow:
            for e in cluster.vision:
                if self.growth[e] is Growth.FULL:      # saturate (would be Growth(3))
                    continue
                self.growth[e] += Growth.INCREMENT
                if self.growth[e] is Growth.FULL:
                    merge_ls.append(e)
        merge_ls.sort(key=self._EDGE_ID.__getitem__)
        for (u, v) in merge_ls:
            cu = self.clusters[self._find(u)]
            cv = self.clusters[self._find(v)]
            cu.vision.remove((u, v))
            cv.vision.discard((u, v))
            self._FORESTER.merge(cu, cv, (u, v))

    def _make_tree(self, tree_root):
        discovered = {tree_root}
        stack = collections.deque([tree_root])
        tree = []
        while stack:
            u = stack.pop()
            for e in sorted(self.erasure.intersection(self.CODE.INCIDENT_EDGES[u]),
                            key=self._EDGE_ID.__getitem__):
                v = self.CODE.traverse_edge(e, u)
                if v not in discovered:
                    discovered.add(v)
                    stack.append(v)
                    tree.append((e, v))
        return tree


from localuf.decoders.node_uf import NodeUF as _RefNodeUF  # noqa: E402


class CanonicalNodeUF(_RefNodeUF):
    """The same pinning for ``NodeUF`` (``localuf/decoders/node_uf.py``): every frontier node of
    an active cluster grows its own active edges (``:45-67``), the round's merge list is
    processed in ascending edge id, the peeling tree is traversed in ascending edge id."""

    def __init__(self, code, dynamic=False, inclination='default'):
        self._EDGE_ID = {e: k for k, e in enumerate(code.EDGES)}
        self.round_count = 0
        super().__init__(code, dynamic=dynamic, inclination=inclination)

    def reset(self, no_boundaries=False):
        super().reset()
        self.round_count = 0

    def _growth_round(self, log_history, clusters_to_grow=None):
        if clusters_to_grow is None:
            clusters_to_grow = self.active_clusters
        self.round_count += 1
        merge_ls, changed = [], set()
        for cluster in clusters_to_grow:
            self._grow(cluster, merge_ls, changed)          # node_uf.py:45-67, unchanged
        merge_ls.sort(key=self._EDGE_ID.__getitem__)
        for (u, v) in merge_ls:
            cu = self.clusters[self._find(u)]
            cv = self.clusters[self._find(v)]
            self._FORESTER.merge(cu, cv, (u, v))

    _make_tree = CanonicalUF._make_tree


from localuf.decoders.node_buf import NodeBUF as _RefNodeBUF  # noqa: E402


class CanonicalNodeBUF(_RefNodeBUF):
    """``NodeBUF`` (``localuf/decoders/node_buf.py``: always the smallest active clusters first)
    with the same two pins; buckets, ``validate`` and ``_update_self_after_union`` are inherited."""

    def __init__(self, code, dynamic=False, inclination='default'):
        self._EDGE_ID = {e: k for k, e in enumerate(code.EDGES)}
        self.round_count = 0
        super().__init__(code, dynamic=dynamic, inclination=inclination)

    def reset(self, no_boundaries=False):
        super().reset()
        self.round_count = 0

    def _growth_round(self, log_history, clusters_to_grow=None):
        self.round_count += 1
        merge_ls, changed = [], set()
        for cluster in list(clusters_to_grow):
            self._grow(cluster, merge_ls, changed)          # node_uf.py:45-67, unchanged
        merge_ls.sort(key=self._EDGE_ID.__getitem__)
        for (u, v) in merge_ls:
            cu = self.clusters[self._find(u)]
            cv = self.clusters[self._find(v)]
            self._FORESTER.merge(cu, cv, (u, v))

    _make_tree = CanonicalUF._make_tree


from localuf.decoders.buf import BUF as _RefBUF  # noqa: E402


class CanonicalBUF(_RefBUF):
    """``BUF`` (``localuf/decoders/buf.py``: always the active clusters with the shortest vision
    first) with the same two pins; buckets, ``validate``, the merges and their bucket
    bookkeeping are inherited."""

    def __init__(self, code, dynamic=False, inclination='default'):
        self._EDGE_ID = {e: k for k, e in enumerate(code.EDGES)}
        self.round_count = 0
        super().__init__(code, dynamic=dynamic, inclination=inclination)

    def reset(self, no_boundaries=False):
        super().reset()
        self.round_count = 0

    def _growth_round(self, log_history, clusters_to_grow=None):
        CanonicalUF._growth_round(self, log_history, list(clusters_to_grow))

    _make_tree = CanonicalUF._make_tree
