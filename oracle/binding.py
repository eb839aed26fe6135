"""ctypes binding of the CPU oracle (oracle/luf_oracle.c).  TEST INFRASTRUCTURE ONLY."""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB = os.path.join(HERE, '_build', 'libluf_oracle.so')

P = C.POINTER


class OracleGraph(C.Structure):
    _fields_ = [
        ('n_nodes', C.c_int32), ('n_boundary', C.c_int32), ('n_edges', C.c_int32),
        ('d', C.c_int32), ('height', C.c_int32),
        ('edge_u', C.c_void_p), ('edge_v', C.c_void_p),
        ('node_flags', C.c_void_p), ('node_id', C.c_void_p), ('node_long', C.c_void_p),
        ('inc_off', C.c_void_p), ('inc_edge', C.c_void_p), ('inc_other', C.c_void_p),
        ('west_edge_mask', C.c_void_p), ('edge_class', C.c_void_p), ('n_classes', C.c_int32),
        ('n_dirs', C.c_int32), ('nbr_edge', C.c_void_p), ('nbr_node', C.c_void_p),
        ('nbr_owned', C.c_void_p),
        ('actis_span', C.c_void_p), ('actis_relay', C.c_void_p), ('actis_global_span', C.c_int32),
        ('noise_kind', C.c_int32),
    ]


def build(force: bool = False) -> str:
    srcs = [os.path.join(HERE, f) for f in os.listdir(HERE) if f.endswith(('.c', '.h')) or f == 'Makefile']
    stale = (not os.path.exists(LIB)) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in srcs)
    if force or stale:
        subprocess.run(['make', '-C', HERE, '-B'], check=True, capture_output=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_uf_decode.restype = C.c_int
        _lib.oracle_decode_batch.restype = C.c_int
        _lib.oracle_logical_batch.restype = C.c_int
        _lib.oracle_mc_batch.restype = C.c_int64
        _lib.oracle_mc_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p,
                                         C.c_uint64, C.c_int64, C.c_int]
        _lib.oracle_decode_batch.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64,
                                             C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p]
    return _lib


class OracleForward(C.Structure):
    _fields_ = [('c', C.c_int32), ('lower_edge', C.c_void_p), ('commit_mask', C.c_void_p), ('art_node', C.c_void_p),
                ('node_lower', C.c_void_p), ('node_side', C.c_void_p)]


class ForwardOracle:
    """Forward-scheme oracle bound to one code (flat graph + localuf_b200._window.ForwardTables)."""

    def __init__(self, code, tables):
        self.orc = Oracle.from_code(code)
        self.t = tables
        self._keep = {}
        f = OracleForward()
        f.c = tables.c
        for name, dt in (('lower_edge', np.int32), ('commit_mask', np.uint32), ('art_node', np.int32),
                         ('node_lower', np.int32), ('node_side', np.uint8)):
            a = np.ascontiguousarray(getattr(tables, name), dtype=dt)
            self._keep[name] = a
            setattr(f, name, a.ctypes.data)
        self.f = f

    def stream(self, decoder, flags, init_bits, fresh_bits):
        WE = self.orc.WE
        fresh = np.ascontiguousarray(fresh_bits, dtype=np.uint32).reshape(-1, WE)
        init = np.ascontiguousarray(init_bits, dtype=np.uint32).reshape(WE)
        n = fresh.shape[0]
        steps = np.zeros((n, 2), np.int32); commit = np.zeros((n, WE), np.uint32); m = np.zeros(n, np.int32)
        lib().oracle_forward_stream.restype = C.c_int
        rc = lib().oracle_forward_stream(C.byref(self.orc.g), C.byref(self.f), decoder, flags, C.c_int32(n),
                                         C.c_void_p(init.ctypes.data), C.c_void_p(fresh.ctypes.data),
                                         C.c_void_p(steps.ctypes.data), C.c_void_p(commit.ctypes.data),
                                         C.c_void_p(m.ctypes.data))
        if rc != 0:
            raise RuntimeError('oracle_forward_stream failed')
        return steps, commit, m


class OracleWindow(C.Structure):
    _fields_ = [
        ('d', C.c_int32), ('h', C.c_int32), ('NLb', C.c_int32), ('NLd', C.c_int32), ('EL', C.c_int32), ('K', C.c_int32),
        ('nb_valid', C.c_void_p), ('nb_q', C.c_void_p), ('nb_dt', C.c_void_p), ('nb_el', C.c_void_p), ('nb_et', C.c_void_p),
        ('edge_q', C.c_void_p), ('edge_dt', C.c_void_p), ('edge_class', C.c_void_p), ('edge_fb', C.c_void_p),
        ('pos_long', C.c_void_p), ('n_classes', C.c_int32), ('merger_slow', C.c_int32), ('schedule_11', C.c_int32),
    ]


class StreamOracle:
    """Snowflake + frugal oracle bound to one window (localuf_b200._window.WindowTables)."""

    def __init__(self, tables, merger='fast', schedule='2:1'):
        self.t = tables
        self._keep = {}
        w = OracleWindow()
        for f in ('d', 'h', 'NLb', 'NLd', 'EL', 'K', 'n_classes'):
            setattr(w, f, int(getattr(tables, f)))
        for name, dt in (('nb_valid', np.uint8), ('nb_q', np.int32), ('nb_dt', np.int32), ('nb_el', np.int32),
                         ('nb_et', np.int32), ('edge_q', np.int32), ('edge_dt', np.int32),
                         ('edge_class', np.uint8), ('edge_fb', np.uint8), ('pos_long', np.int32)):
            a = np.ascontiguousarray(getattr(tables, name), dtype=dt)
            self._keep[name] = a
            setattr(w, name, a.ctypes.data)
        w.merger_slow = int(merger == 'slow')
        w.schedule_11 = int(schedule == '1:1')
        self.w = w
        self.W = (tables.EL + 31) // 32

    def stream(self, fresh_local_bits):
        fresh = np.ascontiguousarray(fresh_local_bits, dtype=np.uint32).reshape(-1, self.W)
        n = fresh.shape[0]
        rt_all = np.zeros(n, np.int32); rt_unr = np.zeros(n, np.int32); m = np.zeros(n, np.int32)
        floor = np.zeros((n, self.W), np.uint32)
        lib().oracle_sf_stream.restype = C.c_int
        rc = lib().oracle_sf_stream(C.byref(self.w), C.c_int32(n), C.c_void_p(fresh.ctypes.data),
                                    C.c_void_p(rt_all.ctypes.data), C.c_void_p(rt_unr.ctypes.data),
                                    C.c_void_p(floor.ctypes.data), C.c_void_p(m.ctypes.data))
        if rc != 0:
            raise RuntimeError('oracle_sf_stream failed')
        return rt_all, rt_unr, floor, m

    def mc(self, class_p, seed, nshots, n, nthreads=1):
        cp = np.ascontiguousarray(class_p, dtype=np.float64)
        cycles = C.c_int64(0)
        lib().oracle_sf_mc.restype = C.c_int64
        m = lib().oracle_sf_mc(C.byref(self.w), C.c_void_p(cp.ctypes.data), C.c_uint64(seed), C.c_int64(nshots),
                               C.c_int32(n), C.c_int(nthreads), C.byref(cycles))
        return int(m), int(cycles.value)


def code_bits_to_local(tables, bits, block):
    """[n, W(E)] bits over code edge ids -> [n, W(EL)] bits over the local index of layer block `block`
    (block 0 = freshly sampled sheet, block h-1 = floor sheet); asserts nothing lies outside the block."""
    bits = np.ascontiguousarray(bits, dtype=np.uint32)
    E = tables.layered_of_code.shape[0]
    u = np.unpackbits(bits.view(np.uint8), axis=1, bitorder='little')[:, :E]
    lay = tables.layered_of_code
    inside = (lay >= block * tables.EL) & (lay < (block + 1) * tables.EL)
    assert not u[:, ~inside].any()
    out = np.zeros((bits.shape[0], 32 * ((tables.EL + 31) // 32)), np.uint8)
    out[:, lay[inside] - block * tables.EL] = u[:, inside]
    return np.packbits(out, axis=1, bitorder='little').view(np.uint32)


NOISE_KIND = {'code capacity': 0, 'phenomenological': 1, 'circuit-level': 2}


class Oracle:
    """Oracle bound to one flat graph (anything with the FlatGraph fields)."""

    def __init__(self, flat, noise_kind: int = 0):
        self.flat = flat
        self._keep = {}
        g = OracleGraph()
        g.n_nodes, g.n_boundary, g.n_edges = flat.n_nodes, flat.n_boundary, flat.n_edges
        g.d, g.height = flat.d, flat.height
        for name, dt in (('edge_u', np.int32), ('edge_v', np.int32), ('node_flags', np.uint8),
                         ('node_id', np.int32), ('node_long', np.int32),
                         ('inc_off', np.int32), ('inc_edge', np.int32), ('inc_other', np.int32),
                         ('west_edge_mask', np.uint32), ('edge_class', np.uint8),
                         ('nbr_edge', np.int32), ('nbr_node', np.int32), ('nbr_owned', np.uint8),
                         ('actis_span', np.int32), ('actis_relay', np.int32)):
            a = np.ascontiguousarray(getattr(flat, name), dtype=dt)
            self._keep[name] = a
            setattr(g, name, a.ctypes.data)
        g.n_classes = flat.n_classes
        g.n_dirs = flat.nbr_edge.shape[1]
        g.noise_kind = noise_kind
        g.actis_global_span = flat.actis_global_span
        self.g = g
        self.WE = (flat.n_edges + 31) // 32
        self.WD = (flat.n_detectors + 31) // 32

    @classmethod
    def from_code(cls, code):
        return cls(code.FLAT, NOISE_KIND[str(code.NOISE)])

    def decode_batch(self, decoder: int, flags: int, syn_bits, want_growth=True):
        syn = np.ascontiguousarray(syn_bits, dtype=np.uint32).reshape(-1, max(self.WD, 1) if self.WD else 0)
        n = syn.shape[0]
        corr = np.zeros((n, self.WE), np.uint32)
        growth = np.zeros((n, self.flat.n_edges), np.uint8) if want_growth else None
        steps = np.zeros((n, 2), np.int32)
        rc = lib().oracle_decode_batch(C.byref(self.g), decoder, flags, n, syn.ctypes.data, corr.ctypes.data,
                                       growth.ctypes.data if want_growth else None, steps.ctypes.data)
        if rc != 0:
            raise RuntimeError(f'oracle_decode_batch failed ({rc})')
        return corr, growth, steps

    def syndrome(self, err_bits):
        err = np.ascontiguousarray(err_bits, dtype=np.uint32).reshape(-1, self.WE)
        syn = np.zeros((err.shape[0], self.WD), np.uint32)
        for s in range(err.shape[0]):
            lib().oracle_syndrome(C.byref(self.g), err[s].ctypes.data_as(C.c_void_p),
                                  syn[s].ctypes.data_as(C.c_void_p))
        return syn

    def logical_batch(self, err_bits, corr_bits):
        err = np.ascontiguousarray(err_bits, dtype=np.uint32).reshape(-1, self.WE)
        corr = np.ascontiguousarray(corr_bits, dtype=np.uint32).reshape(-1, self.WE)
        return np.array([lib().oracle_logical_batch(C.byref(self.g), err[s].ctypes.data_as(C.c_void_p),
                                                    corr[s].ctypes.data_as(C.c_void_p))
                         for s in range(err.shape[0])], np.uint8)

    def sample(self, class_p, seed: int, nshots: int):
        rng = (C.c_uint64 * 4)(*[(seed * 0x9E3779B97F4A7C15 + k * 0xD1B54A32D192ED03 + 1) & (2**64 - 1) for k in range(4)])
        cp = np.ascontiguousarray(class_p, dtype=np.float64)
        err = np.zeros((nshots, self.WE), np.uint32)
        for s in range(nshots):
            lib().oracle_sample(C.byref(self.g), cp.ctypes.data_as(C.c_void_p), rng,
                                err[s].ctypes.data_as(C.c_void_p))
        return err

    def mc_batch(self, decoder: int, flags: int, class_p, seed: int, nshots: int, nthreads: int = 1) -> int:
        cp = np.ascontiguousarray(class_p, dtype=np.float64)
        m = lib().oracle_mc_batch(C.addressof(self.g), decoder, flags, cp.ctypes.data, seed, nshots, nthreads)
        if m < 0:
            raise RuntimeError('oracle_mc_batch failed')
        return int(m)


def dcs_batch(tables, growth_codes, want=('swim', 'uef', 'clustered')):
    """oracle/luf_oracle_dcs.c on [n, E'] growth codes over a localuf_b200._dcs.DcsTables-like object
    -> dict of arrays."""
    L = lib()
    g = np.ascontiguousarray(growth_codes, dtype=np.uint8)
    n, E = g.shape
    assert E == len(tables.edge_u)
    eu = np.ascontiguousarray(tables.edge_u, np.int32); ev = np.ascontiguousarray(tables.edge_v, np.int32)
    w = np.ascontiguousarray(tables.weight, np.float64); ing = np.ascontiguousarray(tables.in_growth, np.uint8)
    side = np.ascontiguousarray(tables.node_side, np.uint8)
    out = {k: np.zeros(n, np.int32 if k == 'clustered' else np.float64) for k in want}
    ptr = lambda k: out[k].ctypes.data if k in out else None
    L.oracle_dcs_batch.restype = C.c_int
    L.oracle_dcs_batch.argtypes = [C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_longlong] + [C.c_void_p] * 4
    rc = L.oracle_dcs_batch(int(tables.n_nodes), E, eu.ctypes.data, ev.ctypes.data, w.ctypes.data, ing.ctypes.data,
                            side.ctypes.data, n, g.ctypes.data, ptr('swim'), ptr('uef'), ptr('clustered'))
    assert rc == 0
    return out
