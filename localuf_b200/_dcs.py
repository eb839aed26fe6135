"""Decoder confidence scores on the device: flat tables for ``luf_dcs_create`` and the ctypes
binding of ``luf_dcs_run`` (``include/luf_b200.h``; kernel ``csrc/luf_dcs.cuh``).

Mirrors ``BaseUF.unclustered_edge_fraction`` / ``swim_distance``
(``localuf/decoders/_base_uf.py:101-121,250-316``) and ``UF.unclustered_node_fraction``
(``localuf/decoders/uf.py:188-197``).  The tables below hold no algorithm: which nodes touch the two meta
boundaries (``_cached_swim_graph``, ``:305-316``), each edge's weight
(``Noise.get_edge_weights``, ``noise/main.py:142-150,332-345``) and whether the edge is a key of
``decoder.growth``.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from localuf_b200 import _engine


@dataclass
class DcsTables:
    n_nodes: int
    edge_u: np.ndarray          # int32[E']
    edge_v: np.ndarray          # int32[E']
    weight: np.ndarray          # float64[E']
    in_growth: np.ndarray       # uint8[E']
    node_side: np.ndarray       # uint8[N]   1 west, 2 east
    growth_words: int           # valid words of a growth row (edges beyond are ungrown)

    @property
    def n_edges(self):
        return len(self.edge_u)


def _weights(code, noise_level) -> np.ndarray:
    table = code.NOISE.get_edge_weights(noise_level)
    return np.array([table[e][1] for e in code.EDGES], np.float64)


def _sides(code) -> np.ndarray:
    flat = code.FLAT
    return np.where(flat.node_long == -1, 1, np.where(flat.node_long == code.D - 1, 2, 0)).astype(np.uint8)


def flat_tables(code, noise_level=None) -> DcsTables:
    """Batch decoders: edge ids and node indices of ``code.FLAT``; every edge is a key of ``growth``."""
    flat = code.FLAT
    E = flat.n_edges
    return DcsTables(n_nodes=flat.n_nodes, edge_u=np.asarray(flat.edge_u, np.int32), edge_v=np.asarray(flat.edge_v, np.int32),
                     weight=_weights(code, noise_level), in_growth=np.ones(E, np.uint8), node_side=_sides(code),
                     growth_words=(E + 15) // 16)


def window_tables(code, window, noise_level=None) -> DcsTables:
    """Snowflake: edge ids are the layered ids of the resident window (block ``b`` from the top at
    ``b * 32 * W(EL)``, ``localuf_b200/_window.py``), nodes those of ``code.FLAT`` (future boundary
    included: its edges are part of the swim graph but not of ``decoder.growth``,
    ``snowflake/main.py:82-85``, ``_base_uf.py:294-298``)."""
    flat = code.FLAT
    t = window
    ELp = 32 * ((t.EL + 31) // 32)
    n = t.h * ELp
    eu, ev = np.zeros(n, np.int32), np.zeros(n, np.int32)
    w, ing = np.zeros(n, np.float64), np.zeros(n, np.uint8)
    weights = _weights(code, noise_level)
    ta, h = code.TIME_AXIS, t.h
    for idx in range(t.h * t.EL):
        b, l = divmod(idx, t.EL)
        k = int(t.code_edge[idx])
        e = code.EDGES[k]
        j = b * ELp + l
        eu[j], ev[j] = flat.edge_u[k], flat.edge_v[k]
        w[j] = weights[k]
        ing[j] = 1 if all(v[ta] < h for v in e) else 0
    return DcsTables(n_nodes=flat.n_nodes, edge_u=eu, edge_v=ev, weight=w, in_growth=ing, node_side=_sides(code),
                     growth_words=n // 16)


class DcsDesc(C.Structure):
    """``luf_dcs_desc`` of include/luf_b200.h."""
    _fields_ = [('n_nodes', C.c_int32), ('n_edges', C.c_int32), ('edge_u', C.c_void_p), ('edge_v', C.c_void_p),
                ('weight', C.c_void_p), ('in_growth', C.c_void_p), ('node_side', C.c_void_p)]


_bound = False


def _lib():
    global _bound
    lib = _engine.library()
    if not _bound:
        v, i64 = C.c_void_p, C.c_int64
        lib.luf_dcs_create.argtypes = [C.POINTER(DcsDesc), C.c_int, C.POINTER(C.c_void_p)]
        lib.luf_dcs_destroy.argtypes = [v]
        lib.luf_dcs_run.argtypes = [v, i64, v, i64, C.c_int32, v, v, v, v]
        lib.luf_dcs_run_host.argtypes = [v, i64, v, i64, C.c_int32, v, v, v]
        _bound = True
    return lib


class DcsEngine:
    """One uploaded swim graph (``luf_dcs``) on one CUDA device.  No CPU path."""

    def __init__(self, tables: DcsTables, device: int = 0):
        lib = _lib()
        if lib.luf_device_count() <= 0:
            raise _engine.EngineUnavailable("no CUDA device visible; the engine has no CPU fallback")
        self.tables = tables
        self.device = device
        d = DcsDesc()
        d.n_nodes, d.n_edges = tables.n_nodes, tables.n_edges
        self._keep = {}
        for name, dt in (('edge_u', np.int32), ('edge_v', np.int32), ('weight', np.float64),
                         ('in_growth', np.uint8), ('node_side', np.uint8)):
            a = np.ascontiguousarray(getattr(tables, name), dtype=dt)
            self._keep[name] = a
            setattr(d, name, a.ctypes.data)
        h = C.c_void_p()
        _engine._check(lib.luf_dcs_create(C.byref(d), device, C.byref(h)))
        self._h = h

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        lib = getattr(_engine, '_lib', None) if _engine is not None else None      # None at interpreter shutdown
        if h and lib is not None:
            lib.luf_dcs_destroy(h)

    def run_host(self, growth2b, want_swim=True, want_uef=True, want_clustered=True):
        """growth2b uint32 [n, words] -> (swim float64[n] | None, uef float64[n] | None, clustered int32[n] | None)."""
        g = np.ascontiguousarray(growth2b, dtype=np.uint32)
        g = g.reshape(-1, g.shape[-1])
        n, stride = g.shape
        swim = np.zeros(n, np.float64) if want_swim else None
        uef = np.zeros(n, np.float64) if want_uef else None
        cl = np.zeros(n, np.int32) if want_clustered else None
        _engine._check(_lib().luf_dcs_run_host(self._h, n, g.ctypes.data, stride, min(stride, self.tables.growth_words),
                                               _engine._ptr(swim), _engine._ptr(uef), _engine._ptr(cl)))
        return swim, uef, cl

    def run(self, nshots, growth2b, stride_words, swim=None, uef=None, clustered=None):
        """Device buffers (torch tensors), current torch stream."""
        _engine._check(_lib().luf_dcs_run(self._h, nshots, _engine._ptr(growth2b), stride_words,
                                          min(stride_words, self.tables.growth_words), _engine._ptr(swim),
                                          _engine._ptr(uef), _engine._ptr(clustered), _engine._stream()))


def pack_growth(codes: np.ndarray) -> np.ndarray:
    """[n, E] uint8 growth codes -> [n, ceil(E / 16)] words of 2-bit fields (inverse of ``_engine.unpack_growth``)."""
    c = np.ascontiguousarray(codes, np.uint8)
    n, E = c.shape
    W = (E + 15) // 16
    pad = np.zeros((n, W * 16), np.uint32)
    pad[:, :E] = c
    sh = (2 * np.arange(16, dtype=np.uint32))[None, None, :]
    return (pad.reshape(n, W, 16) << sh).sum(axis=2, dtype=np.uint64).astype(np.uint32)
