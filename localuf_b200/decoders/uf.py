"""Union-Find decoder (Delfosse & Nickerson), device-backed.

Same constructor, attributes and error behaviour as the reference ``UF``
(``localuf/decoders/uf.py:16-119``); the work happens in the CUDA kernels of
``csrc/luf_core.cuh`` (grow / merge / peel in shared memory, one warp per
shot).  Where the reference leaves the order of a round's merges and of the
peeling traversal to CPython's ``set`` layout, the kernels use ascending edge
id -- one legal execution of the reference algorithm ("UF parity
contract").
"""
from __future__ import annotations

from localuf_b200 import _engine
from localuf_b200.decoders._base import BaseUF
from localuf_b200.schemes import Frugal


class UF(BaseUF):
    _DECODER_ID = _engine.DEC_UF

    def __init__(self, code, dynamic: bool = False, inclination: str = 'default'):
        if isinstance(code.SCHEME, Frugal):
            raise ValueError('UF incompatible with frugal scheme.')
        if inclination not in ('default', 'west'):
            raise ValueError(f"Unknown inclination {inclination!r}.")
        super().__init__(code)
        self._flags = (_engine.UF_DYNAMIC if dynamic else 0) | (_engine.UF_WEST if inclination == 'west' else 0)
        self._pending: set = set()
        self.round_count = 0

    def __repr__(self):
        return f'UF(code={self.CODE}, syndrome={self.syndrome})'

    def decode(self, syndrome, draw=False, fig_width=None):
        """``uf.py:106-119``: validate then peel; result in ``self.correction``; returns None."""
        if draw:
            raise NotImplementedError("Drawing is outside this engine (SURVEY.md section 2, row 14).")
        self.validate(syndrome)
        self.peel()

    def validate(self, syndrome, log_history=False):
        """``uf.py:200-209``.  The device decodes in one go; ``peel`` publishes the correction."""
        self.syndrome = set(syndrome)
        self._pending, steps = self._decode_one(self.syndrome)
        self.round_count = int(steps[0])

    def peel(self):
        """``uf.py:305-312``."""
        self.correction = set(self._pending)

    def reset(self, no_boundaries=False):
        if no_boundaries:
            raise NotImplementedError("no_boundaries (complementary gap) is outside the hot path.")
        super().reset()
        self._pending = set()
        self.round_count = 0

    def unclustered_node_fraction(self) -> float:
        """``uf.py:188-197``: the fraction of nodes in no cluster of more than one node (the nodes
        without a fully grown edge -- every such edge has been merged)."""
        from localuf_b200 import _dcs
        _, _, cl = self._dcs_engine(None).run_host(_dcs.pack_growth(self._growth_codes[None, :]),
                                                   want_swim=False, want_uef=False)
        return 1 - int(cl[0]) / self._CODE.NODE_COUNT

    def complementary_gap(self, noise_level=None, **kwargs):
        """``uf.py:121-177``: needs ``merge_equivalent_boundary_nodes=True`` graphs, which this engine builds
        but does not decode (out of scope)."""
        if not getattr(self._CODE, '_MERGED_EQUIVALENT_BOUNDARY_NODES', False):
            raise ValueError('Complementary gap requires all equivalent boundary nodes in the decoding graph be merged.')
        raise NotImplementedError("complementary_gap is outside the hot path (DESIGN.md section 8).")

    # -------------------------------------------------------- batched entry
    def _mc_batch(self, noise_level: float, n: int, seed=None) -> int:
        """Failure count of ``n`` fused sample->decode->check shots (``_schemes.py:116-155``)."""
        eng = self.engine
        return self._sharded_counts(
            lambda s, lo, cnt: eng.mc_run_host(self._DECODER_ID, self._flags, noise_level, s, lo, cnt), n, seed)


class NodeUF(UF):
    """``localuf/decoders/node_uf.py:5-94``: the original Union-Find growth -- every frontier
    node of an active cluster grows each of its active edges, so an edge inside one cluster
    grows from both ends (the growth rule Macar, Actis and Snowflake use too).  Same kernels as
    :class:`UF` with the ``LUF_UF_NODE`` flag."""

    def __init__(self, code, dynamic: bool = False, inclination: str = 'default'):
        super().__init__(code, dynamic=dynamic, inclination=inclination)
        self._flags |= _engine.UF_NODE

    def __repr__(self):
        return f'NodeUF(code={self.CODE}, syndrome={self.syndrome})'


class NodeBUF(NodeUF):
    """``localuf/decoders/node_buf.py:5-61`` (the bucket UF of Huang et al. 2020): in every
    growth round only the active clusters of the smallest size grow.  Same kernels with
    ``LUF_UF_NODE | LUF_UF_BUCKET``.  The reference sizes its bucket table for ``(d+1)*d`` nodes
    and raises ``KeyError`` when a cluster outgrows it (any graph with measurement rounds, near
    threshold); the device has no such limit."""

    def __init__(self, code, dynamic: bool = False, inclination: str = 'default'):
        super().__init__(code, dynamic=dynamic, inclination=inclination)
        self._flags |= _engine.UF_BUCKET

    def __repr__(self):
        return f'NodeBUF(code={self.CODE}, syndrome={self.syndrome})'


class BUF(UF):
    """``localuf/decoders/buf.py:6-116`` (bucket UF): in every growth round only the active
    clusters with the shortest vision -- the fewest not-yet-fully-grown edges around them --
    grow.  Same kernels with ``LUF_UF_BUCKET``.  Where the reference raises ``ValueError``
    (``Growth(3)``: an edge one cluster grew to HALF is later grown by both its clusters in one
    round) growth saturates at FULL here."""

    def __init__(self, code, dynamic: bool = False, inclination: str = 'default'):
        super().__init__(code, dynamic=dynamic, inclination=inclination)
        self._flags |= _engine.UF_BUCKET

    def __repr__(self):
        return f'BUF(code={self.CODE}, syndrome={self.syndrome})'
