"""Build + bind the host emulation of the kernel phases (tests/emu/emu.cpp).
TEST INFRASTRUCTURE ONLY -- checks kernel *logic* on the CPU suite."""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
LIB = os.environ.get('LUF_EMU_LIB') or os.path.join(HERE, '_build', 'libluf_emu.so')
CSRC = os.path.join(ROOT, 'localuf_b200', 'csrc')
SRCS = [os.path.join(HERE, 'emu.cpp')] + [
    os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(('.cuh', '.h', '.inl'))
] + [os.path.join(ROOT, 'include', 'luf_b200.h')]


def build(tiny_lists=False):
    """``tiny_lists``: a second build whose in-kernel work lists hold 3 entries, so that
    every overflow path of the kernel phases runs on ordinary test inputs."""
    if os.environ.get('LUF_EMU_LIB') and not tiny_lists:
        return LIB                      # a prebuilt variant (tests/test_emu_sanitized.py)
    out = LIB.replace('.so', '_tiny.so') if tiny_lists else LIB
    if (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in SRCS):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(['g++', '-O1', '-g', '-fPIC', '-shared', '-std=c++17', '-Wall', '-Wextra']
                       + (['-DLUF_LIST_CAP=3', '-DLUF_LIST2_CAP=2', '-DLUF_CTA_LIST_CAP=5', '-DLUF_CTA_LIST2_CAP=3', '-DLUF_CA_TLIST_CAP=4', '-DLUF_SF_ALIST_CAP=4', '-DLUF_SF_CTA_ALIST_CAP=6', '-DLUF_CA_CTA_TLIST_CAP=5', '-DLUF_WIDE_LIST_CAP=5', '-DLUF_WIDE_LIST2_CAP=3'] if tiny_lists else []) + ['-o', out, SRCS[0]],
                       check=True, capture_output=True)
    return out


_lib = None
_libs = {}


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
    return _lib


def build_sanitized():
    """The emulation with AddressSanitizer + UBSan (heap buffers = the shot's shared-memory
    slice: an out-of-range index of the kernel phases is reported).  Returns (library, libasan)."""
    out = LIB.replace('.so', '_asan.so')
    if (not os.path.exists(out)) or any(os.path.getmtime(s) > os.path.getmtime(out) for s in SRCS):
        os.makedirs(os.path.dirname(out), exist_ok=True)
        subprocess.run(['g++', '-O1', '-g', '-fPIC', '-shared', '-std=c++17', '-fsanitize=address,undefined',
                        '-fno-sanitize-recover=all', '-o', out, SRCS[0]], check=True, capture_output=True)
    asan = subprocess.run(['g++', '-print-file-name=libasan.so'], capture_output=True, text=True).stdout.strip()
    return out, asan


def use_tiny_lists(on: bool):
    """Switch the library the bindings call (tests only)."""
    global _lib
    _lib = _libs.setdefault(on, C.CDLL(build(tiny_lists=on)))


def _vp(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


class Emu:
    def __init__(self, code):
        from localuf_b200._engine import make_desc, NOISE_KIND
        self.flat = code.FLAT
        self.noise = code.NOISE
        self.desc, self._keep = make_desc(self.flat, NOISE_KIND[str(code.NOISE)])
        self.WE = (self.flat.n_edges + 31) // 32
        self.WD = (self.flat.n_detectors + 31) // 32
        self.WG = (self.flat.n_edges + 15) // 16

    def decode_uf(self, flags, syn_bits):
        syn = np.ascontiguousarray(syn_bits, dtype=np.uint32).reshape(-1, self.WD)
        n = syn.shape[0]
        corr = np.zeros((n, self.WE), np.uint32)
        growth = np.zeros((n, self.WG), np.uint32)
        steps = np.zeros((n, 2), np.int32)
        rc = lib().emu_decode_uf(C.byref(self.desc), flags, C.c_longlong(n), _vp(syn),
                                 _vp(corr), _vp(growth), _vp(steps))
        assert rc == 0
        return corr, growth, steps

    def decode_uf_cta(self, flags, syn_bits, nl=128):
        """The block-per-shot instance of the phases (luf::uf_cta) with a tile of ``nl`` lanes."""
        syn = np.ascontiguousarray(syn_bits, dtype=np.uint32).reshape(-1, self.WD)
        n = syn.shape[0]
        corr = np.zeros((n, self.WE), np.uint32)
        growth = np.zeros((n, self.WG), np.uint32)
        steps = np.zeros((n, 2), np.int32)
        rc = lib().emu_decode_uf_cta(C.byref(self.desc), flags, nl, C.c_longlong(n), _vp(syn),
                                     _vp(corr), _vp(growth), _vp(steps))
        assert rc == 0
        return corr, growth, steps

    def decode_uf_wide(self, flags, syn_bits, nl=128):
        """The wide instance of the phases (luf::uf_wide: 64-bit node words) with a tile of ``nl`` lanes."""
        syn = np.ascontiguousarray(syn_bits, dtype=np.uint32).reshape(-1, self.WD)
        n = syn.shape[0]
        corr = np.zeros((n, self.WE), np.uint32)
        growth = np.zeros((n, self.WG), np.uint32)
        steps = np.zeros((n, 2), np.int32)
        rc = lib().emu_decode_uf_wide(C.byref(self.desc), flags, nl, C.c_longlong(n), _vp(syn),
                                      _vp(corr), _vp(growth), _vp(steps))
        assert rc == 0
        return corr, growth, steps

    def mc_uf_wide(self, flags, noise_level, seed, shot0, nshots, nl=128, parity_only=False):
        """Fused shot loop of the wide instance: (failures, err, syn, corr, logical)."""
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        counts = (C.c_ulonglong * 2)(0, 0)
        err = np.zeros((nshots, self.WE), np.uint32)
        syn = np.zeros((nshots, self.WD), np.uint32)
        corr = np.zeros((nshots, self.WE), np.uint32)
        logical = np.zeros(nshots, np.uint8)
        rc = lib().emu_mc_uf_wide(C.byref(self.desc), flags, nl, int(parity_only), _vp(cp), C.c_ulonglong(seed),
                                  C.c_ulonglong(shot0), C.c_longlong(nshots), counts,
                                  _vp(err), _vp(syn), _vp(corr), _vp(logical))
        assert rc == 0
        return int(counts[0]), err, syn, corr, logical

    def sample_weight(self, weight, seed, shot0, nshots):
        err = np.zeros((nshots, self.WE), np.uint32)
        syn = np.zeros((nshots, self.WD), np.uint32)
        rc = lib().emu_sample_weight(C.byref(self.desc), int(weight), C.c_ulonglong(seed), C.c_ulonglong(shot0),
                                     C.c_longlong(nshots), _vp(err), _vp(syn))
        assert rc == 0
        return err, syn

    def mc_uf(self, flags, noise_level, seed, shot0, nshots, export=True):
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        counts = (C.c_ulonglong * 2)(0, 0)
        err = np.zeros((nshots, self.WE), np.uint32) if export else None
        syn = np.zeros((nshots, self.WD), np.uint32) if export else None
        corr = np.zeros((nshots, self.WE), np.uint32) if export else None
        rc = lib().emu_mc_uf(C.byref(self.desc), flags, _vp(cp), C.c_ulonglong(seed),
                             C.c_ulonglong(shot0), C.c_longlong(nshots), counts,
                             _vp(err), _vp(syn), _vp(corr))
        assert rc == 0
        return int(counts[0]), err, syn, corr

    def mc_uf_parity(self, flags, noise_level, seed, nshots, nl=32):
        """(failures by the full peel, failures by uf_peel_parity, shots where they differ, shots with a
        cluster that holds boundary nodes of both sides)."""
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        counts = (C.c_ulonglong * 4)(0, 0, 0, 0)
        rc = lib().emu_mc_uf_parity(C.byref(self.desc), flags, _vp(cp), C.c_ulonglong(seed), C.c_ulonglong(0),
                                    C.c_longlong(nshots), counts, nl)
        assert rc == 0, rc
        return tuple(int(x) for x in counts)

    def mc_uf_sides(self, flags, noise_level, seed, nshots, nl=32):
        """Canonical validate, then the west parity by the pass over the defects (no peel, g.leaf_bnd) and by the
        full peel, on every shot: (failures full peel, failures defect pass, shots where they differ, shots with
        a cluster on both sides)."""
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        counts = (C.c_ulonglong * 4)(0, 0, 0, 0)
        rc = lib().emu_mc_uf_sides(C.byref(self.desc), flags, _vp(cp), C.c_ulonglong(seed), C.c_ulonglong(0),
                                   C.c_longlong(nshots), counts, nl)
        assert rc == 0, rc
        return tuple(int(x) for x in counts)

    def parity_ok(self):
        """Whether pack_graph found the graph fit for the fused kernels' parity-only peel."""
        return int(lib().emu_parity_ok(C.byref(self.desc)))

    def mc_uf_fast(self, flags, noise_level, seed, nshots, nl=32):
        """(failures canonical, failures by the order-free merge + parity path, shots where they differ,
        mixed shots, shots whose round counts differ)."""
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        counts = (C.c_ulonglong * 5)(0, 0, 0, 0, 0)
        rc = lib().emu_mc_uf_fast(C.byref(self.desc), flags, _vp(cp), C.c_ulonglong(seed), C.c_ulonglong(0),
                                  C.c_longlong(nshots), counts, nl)
        assert rc == 0, rc
        return tuple(int(x) for x in counts)

    def mc_uf_compact(self, noise_level, seed, shot0, nshots, nl=32, mu=-1.0, want_stats=False, west=False):
        """The compact fast path (luf_uf_fast.cuh): per shot 0 / 1 = logical bit, 2 = HARD (queued for the
        canonical kernel; none with ``west``); stats [nshots, 4] = rounds, touched nodes, most newly FULL edges / active nodes of a round."""
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        out = np.zeros(nshots, np.uint8)
        stats = np.zeros((nshots, 4), np.uint32) if want_stats else None
        lib().emu_mc_uf_compact.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_longlong, C.c_int,
                                            C.c_double, C.c_void_p, C.c_void_p]
        rc = lib().emu_mc_uf_compact(C.addressof(self.desc), int(west), _vp(cp), seed, shot0, nshots, nl, float(mu), _vp(out), _vp(stats))
        assert rc == 0, rc
        return (out, stats) if want_stats else out

    def mc_uf_compact_log(self, noise_level, seed, shot0, nshots, nl=32, mu=-1.0, lcap=-1, want_keys=False, onepass=False):
        """The compact fast path with the merge log and hard_resolve (default inclination): per shot the logical
        bit, or 2 where the log could not hold the shot; optionally the number of exported keys per shot."""
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        out = np.zeros(nshots, np.uint8)
        nkeys = np.zeros(nshots, np.uint32)
        self.last_log_entries = np.zeros(nshots, np.uint32)            # newly FULL edges of each shot (what the log has to hold)
        lib().emu_mc_uf_compact_log.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_longlong, C.c_int,
                                                C.c_double, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        rc = lib().emu_mc_uf_compact_log(C.addressof(self.desc), _vp(cp), seed, shot0, nshots, nl, float(mu), int(lcap), int(onepass), _vp(out), _vp(nkeys),
                                         _vp(self.last_log_entries))
        assert rc == 0, rc
        return (out, nkeys) if want_keys else out

    def sample_compact(self, noise_level, seed, shot0, nshots):
        """The staged sampler on the compact tables (f_sample_record): (err [n, W(E)], syn [n, W(D)])."""
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        err = np.zeros((nshots, self.WE), np.uint32)
        syn = np.zeros((nshots, self.WD), np.uint32)
        lib().emu_sample_compact.argtypes = [C.c_void_p, C.c_void_p, C.c_ulonglong, C.c_ulonglong, C.c_longlong, C.c_void_p, C.c_void_p]
        rc = lib().emu_sample_compact(C.addressof(self.desc), _vp(cp), seed, shot0, nshots, _vp(err), _vp(syn))
        assert rc == 0, rc
        return err, syn

    def decode_ca(self, decoder, flags, syn_bits, nl=32):
        syn = np.ascontiguousarray(syn_bits, dtype=np.uint32).reshape(-1, self.WD)
        n = syn.shape[0]
        corr = np.zeros((n, self.WE), np.uint32)
        growth = np.zeros((n, self.WG), np.uint32)
        steps = np.zeros((n, 2), np.int32)
        rc = lib().emu_decode_ca_tile(C.byref(self.desc), decoder, flags, C.c_longlong(n), _vp(syn),
                                      _vp(corr), _vp(growth), _vp(steps), nl)
        assert rc == 0, rc
        return corr, growth, steps

    def forward_stream(self, code, decoder, flags, init_bits, fresh_bits):
        from localuf_b200 import _window
        from localuf_b200._engine import make_forward_desc
        fdesc, keep = make_forward_desc(_window.build_forward(code))
        fresh = np.ascontiguousarray(fresh_bits, dtype=np.uint32).reshape(-1, self.WE)
        init = np.ascontiguousarray(init_bits, dtype=np.uint32).reshape(self.WE)
        n = fresh.shape[0]
        steps = np.zeros((n, 2), np.int32); commit = np.zeros((n, self.WE), np.uint32); m = np.zeros(n, np.int32)
        rc = lib().emu_fw_stream(C.byref(self.desc), C.byref(fdesc), decoder, flags, n, _vp(init), _vp(fresh),
                                 _vp(steps), _vp(commit), _vp(m))
        assert rc == 0, rc
        return steps, commit, m


    def _forward_stream_wide(self, code, flags, init_bits, fresh_bits, nl=96):
        """The forward scheme on the wide instance (luf_fw_wide.cuh), UF."""
        from localuf_b200 import _window
        from localuf_b200._engine import make_forward_desc
        fdesc, keep = make_forward_desc(_window.build_forward(code))
        fresh = np.ascontiguousarray(fresh_bits, dtype=np.uint32).reshape(-1, self.WE)
        init = np.ascontiguousarray(init_bits, dtype=np.uint32).reshape(self.WE)
        n = fresh.shape[0]
        steps = np.zeros((n, 2), np.int32); commit = np.zeros((n, self.WE), np.uint32); m = np.zeros(n, np.int32)
        rc = lib().emu_fw_stream_wide(C.byref(self.desc), C.byref(fdesc), flags, nl, n, _vp(init), _vp(fresh),
                                      _vp(steps), _vp(commit), _vp(m))
        assert rc == 0, rc
        return steps, commit, m


Emu.forward_stream_wide = Emu._forward_stream_wide


class EmuStream:
    """Host emulation of the Snowflake / frugal kernel phases for one window."""

    def __init__(self, code, tables=None):
        from localuf_b200 import _window
        from localuf_b200._engine import make_window_desc
        self.tables = _window.build(code) if tables is None else tables
        self.noise = code.NOISE
        self.desc, self._keep = make_window_desc(self.tables)
        self.W = (self.tables.EL + 31) // 32

    def stream(self, flags, fresh_local_bits, nl=32):
        """``nl`` = 32: the warp-tile drivers; any other width: the block-tile drivers (luf::sf_cta)."""
        fresh = np.ascontiguousarray(fresh_local_bits, dtype=np.uint32).reshape(-1, self.W)
        n = fresh.shape[0]
        rt = np.zeros((n, 2), np.int32); m = np.zeros(n, np.int32); floor = np.zeros((n, self.W), np.uint32)
        rc = lib().emu_sf_stream_tile(C.byref(self.desc), flags, n, _vp(fresh), _vp(rt), _vp(floor), _vp(m), nl)
        assert rc == 0, rc
        return rt, floor, m

    def stream_metrics(self, flags, fresh_local_bits, nl=32, drop_only=None):
        """As :meth:`stream`, plus the window's growth words and (min_defect_height, min_active_layer) per cycle."""
        fresh = np.ascontiguousarray(fresh_local_bits, dtype=np.uint32).reshape(-1, self.W)
        n = fresh.shape[0]
        rt = np.zeros((n, 2), np.int32); m = np.zeros(n, np.int32); floor = np.zeros((n, self.W), np.uint32)
        growth = np.zeros((n, self.tables.h * 2 * self.W), np.uint32); mins = np.zeros((n, 2), np.int32)
        drop = None if drop_only is None else np.ascontiguousarray(drop_only, dtype=np.uint8)
        rc = lib().emu_sf_stream_metrics(C.byref(self.desc), flags, n, _vp(fresh), _vp(rt), _vp(floor), _vp(m),
                                         _vp(growth), _vp(mins), _vp(drop), nl)
        assert rc == 0, rc
        return rt, floor, m, growth, mins

    def latency(self, flags, noise_level, seed, stream0, nstreams, nl=32):
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        lat = np.zeros((nstreams, 3), np.int32)
        rc = lib().emu_sf_latency(C.byref(self.desc), flags, _vp(cp), C.c_ulonglong(seed), C.c_ulonglong(stream0),
                                  C.c_longlong(nstreams), _vp(lat), nl)
        assert rc == 0, rc
        return lat

    def mc(self, flags, noise_level, seed, stream0, nstreams, n, nl=32):
        cp = np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)
        counts = (C.c_ulonglong * 4)(0, 0, 0, 0)
        steps = np.zeros((nstreams, n, 2), np.int32)
        rc = lib().emu_sf_mc_tile(C.byref(self.desc), flags, _vp(cp), C.c_ulonglong(seed), C.c_ulonglong(stream0),
                                  C.c_longlong(nstreams), n, counts, _vp(steps), nl)
        assert rc == 0, rc
        return [int(x) for x in counts], steps
