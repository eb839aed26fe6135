"""ctypes binding of the CUDA engine (``libluf_b200.so``, C ABI in
``include/luf_b200.h``).

PyTorch is used for plumbing only: device buffers are ``torch`` tensors whose
raw pointers are handed to the C ABI, and the current ``torch`` CUDA stream is
the stream the kernels are launched on.  There is **no CPU fallback**: if the
library or a CUDA device is missing, every entry point raises
:class:`EngineUnavailable`.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
# LUF_B200_LIB: an alternative build of the same sources (tools/variants.py), never a different implementation
LIB_PATH = os.environ.get('LUF_B200_LIB') or os.path.join(HERE, 'csrc', 'libluf_b200.so')

DEC_UF, DEC_MACAR, DEC_ACTIS, DEC_SNOWFLAKE = 0, 1, 2, 3
UF_DYNAMIC, UF_WEST, UF_NODE, UF_BUCKET, ACTIS_UNOPTIMAL = 1, 2, 4, 8, 4
NOISE_KIND = {'code capacity': 0, 'phenomenological': 1, 'circuit-level': 2}


class EngineUnavailable(RuntimeError):
    """The CUDA engine cannot run here (library not built, or no GPU)."""


class GraphDesc(C.Structure):
    """``luf_graph_desc`` of include/luf_b200.h."""
    _fields_ = [
        ('n_nodes', C.c_int32), ('n_boundary', C.c_int32), ('n_edges', C.c_int32),
        ('d', C.c_int32), ('height', C.c_int32), ('noise_kind', C.c_int32),
        ('edge_u', C.c_void_p), ('edge_v', C.c_void_p),
        ('node_flags', C.c_void_p), ('node_id', C.c_void_p), ('node_long', C.c_void_p),
        ('inc_off', C.c_void_p), ('inc_edge', C.c_void_p), ('inc_other', C.c_void_p),
        ('west_edge_mask', C.c_void_p), ('edge_class', C.c_void_p), ('n_classes', C.c_int32),
        ('n_dirs', C.c_int32), ('nbr_edge', C.c_void_p), ('nbr_node', C.c_void_p),
        ('nbr_owned', C.c_void_p),
        ('actis_span', C.c_void_p), ('actis_relay', C.c_void_p), ('actis_global_span', C.c_int32),
    ]


_FIELD_TYPES = (
    ('edge_u', np.int32), ('edge_v', np.int32), ('node_flags', np.uint8), ('node_id', np.int32),
    ('node_long', np.int32), ('inc_off', np.int32), ('inc_edge', np.int32), ('inc_other', np.int32),
    ('west_edge_mask', np.uint32), ('edge_class', np.uint8),
    ('nbr_edge', np.int32), ('nbr_node', np.int32), ('nbr_owned', np.uint8),
    ('actis_span', np.int32), ('actis_relay', np.int32),
)


def make_desc(flat, noise_kind: int):
    """Return ``(GraphDesc, keepalive)`` for a :class:`FlatGraph`."""
    keep = {}
    d = GraphDesc()
    d.n_nodes, d.n_boundary, d.n_edges = flat.n_nodes, flat.n_boundary, flat.n_edges
    d.d, d.height, d.noise_kind = flat.d, flat.height, noise_kind
    for name, dt in _FIELD_TYPES:
        a = np.ascontiguousarray(getattr(flat, name), dtype=dt)
        keep[name] = a
        setattr(d, name, a.ctypes.data)
    d.n_classes = flat.n_classes
    d.n_dirs = flat.nbr_edge.shape[1]
    d.actis_global_span = flat.actis_global_span
    return d, keep


class ForwardDesc(C.Structure):
    """``luf_forward_desc`` of include/luf_b200.h."""
    _fields_ = [('commit_height', C.c_int32), ('n_cleanse', C.c_int32), ('lower_edge', C.c_void_p),
                ('commit_mask', C.c_void_p), ('art_node', C.c_void_p), ('edge_fb', C.c_void_p),
                ('node_lower', C.c_void_p), ('node_side', C.c_void_p), ('buffer_class', C.c_void_p)]


def make_forward_desc(t):
    """Return ``(ForwardDesc, keepalive)`` for a :class:`localuf_b200._window.ForwardTables`."""
    keep = {}
    d = ForwardDesc()
    d.commit_height, d.n_cleanse = t.c, t.n_cleanse
    for name, dt in (('lower_edge', np.int32), ('commit_mask', np.uint32), ('art_node', np.int32),
                     ('edge_fb', np.uint8), ('node_lower', np.int32), ('node_side', np.uint8),
                     ('buffer_class', np.uint8)):
        a = np.ascontiguousarray(getattr(t, name), dtype=dt)
        keep[name] = a
        setattr(d, name, a.ctypes.data)
    return d, keep


class WindowDesc(C.Structure):
    """``luf_window_desc`` of include/luf_b200.h."""
    _fields_ = [
        ('d', C.c_int32), ('h', C.c_int32), ('n_boundary_per_layer', C.c_int32),
        ('n_detectors_per_layer', C.c_int32), ('edges_per_layer', C.c_int32), ('n_dirs', C.c_int32),
        ('nb_valid', C.c_void_p), ('nb_q', C.c_void_p), ('nb_dt', C.c_void_p), ('nb_el', C.c_void_p),
        ('nb_et', C.c_void_p), ('edge_q', C.c_void_p), ('edge_dt', C.c_void_p), ('edge_class', C.c_void_p),
        ('edge_fb', C.c_void_p), ('pos_long', C.c_void_p), ('n_classes', C.c_int32),
    ]


def make_window_desc(t):
    """Return ``(WindowDesc, keepalive)`` for a :class:`localuf_b200._window.WindowTables`."""
    keep = {}
    d = WindowDesc()
    d.d, d.h, d.n_boundary_per_layer, d.n_detectors_per_layer = t.d, t.h, t.NLb, t.NLd
    d.edges_per_layer, d.n_dirs, d.n_classes = t.EL, t.K, t.n_classes
    for name, dt in (('nb_valid', np.uint8), ('nb_q', np.int32), ('nb_dt', np.int32), ('nb_el', np.int32),
                     ('nb_et', np.int32), ('edge_q', np.int32), ('edge_dt', np.int32), ('edge_class', np.uint8),
                     ('edge_fb', np.uint8), ('pos_long', np.int32)):
        a = np.ascontiguousarray(getattr(t, name), dtype=dt)
        keep[name] = a
        setattr(d, name, a.ctypes.data)
    return d, keep


SF_SLOW_MERGER, SF_ONE_ONE, SF_SYNDROME, SF_LATENCY_TAIL = 1, 2, 4, 8

_lib = None
_lib_lock = threading.Lock()


def _load():
    global _lib
    with _lib_lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise EngineUnavailable(
                f"{LIB_PATH} is not built; run `python -c 'import __graft_entry__ as g; g.build()'`. "
                "The engine has no CPU fallback.")
        lib = C.CDLL(LIB_PATH)
        lib.luf_last_error.restype = C.c_char_p
        lib.luf_launch_count.restype = C.c_int64
        lib.luf_launch_count.argtypes = [C.c_void_p]
        lib.luf_graph_create.argtypes = [C.POINTER(GraphDesc), C.c_int, C.POINTER(C.c_void_p)]
        lib.luf_graph_destroy.argtypes = [C.c_void_p]
        lib.luf_graph_words.argtypes = [C.c_void_p, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
        v, u64, i64 = C.c_void_p, C.c_uint64, C.c_int64
        lib.luf_sample.argtypes = [v, v, u64, u64, i64, v, v, v]
        lib.luf_decode_batch.argtypes = [v, C.c_int, C.c_int, i64, v, v, v, v, v]
        lib.luf_check.argtypes = [v, i64, v, v, v, v, v]
        lib.luf_mc_run.argtypes = [v, C.c_int, C.c_int, v, u64, u64, i64, v, v]
        lib.luf_count_strings.argtypes = [v, i64, v, v, v, v, v]
        lib.luf_mc_run_host.argtypes = [v, C.c_int, C.c_int, v, u64, u64, i64, v]
        lib.luf_mc_run_logical.argtypes = [v, C.c_int, C.c_int, v, u64, u64, i64, v, v, v]
        lib.luf_mc_run_logical_host.argtypes = [v, C.c_int, C.c_int, v, u64, u64, i64, v, v]
        lib.luf_decode_batch_host.argtypes = [v, C.c_int, C.c_int, i64, v, v, v, v]
        lib.luf_decode_frames.argtypes = [v, C.c_int, C.c_int, i64, v, v, v, v]
        lib.luf_decode_frames_host.argtypes = [v, C.c_int, C.c_int, i64, v, v, v]
        lib.luf_sample_host.argtypes = [v, v, u64, u64, i64, v, v]
        lib.luf_check_host.argtypes = [v, i64, v, v, v, v]
        lib.luf_sample_weight.argtypes = [v, C.c_int32, u64, u64, i64, v, v, v]
        lib.luf_sample_weight_host.argtypes = [v, C.c_int32, u64, u64, i64, v, v]
        lib.luf_forward_attach.argtypes = [v, C.POINTER(ForwardDesc)]
        lib.luf_forward_decode_host.argtypes = [v, C.c_int, C.c_int, i64, C.c_int32, v, v, v, v, v]
        lib.luf_forward_mc_host.argtypes = [v, C.c_int, C.c_int, v, u64, u64, i64, C.c_int32, v, v]
        lib.luf_forward_decode.argtypes = [v, C.c_int, C.c_int, i64, C.c_int32, v, v, v, v, v, v]
        lib.luf_forward_mc.argtypes = [v, C.c_int, C.c_int, v, u64, u64, i64, C.c_int32, v, v, v]
        lib.luf_stream_create.argtypes = [C.POINTER(WindowDesc), C.c_int, C.POINTER(C.c_void_p)]
        lib.luf_stream_destroy.argtypes = [v]
        lib.luf_stream_launch_count.restype = C.c_int64
        lib.luf_stream_launch_count.argtypes = [v]
        lib.luf_stream_decode.argtypes = [v, C.c_int, i64, C.c_int32, v, v, v, v, v]
        lib.luf_stream_mc.argtypes = [v, C.c_int, v, u64, u64, i64, C.c_int32, v, v, v]
        lib.luf_stream_decode_host.argtypes = [v, C.c_int, i64, C.c_int32, v, v, v, v]
        lib.luf_stream_mc_host.argtypes = [v, C.c_int, v, u64, u64, i64, C.c_int32, v, v]
        lib.luf_stream_growth_words.argtypes = [v]
        lib.luf_stream_decode_metrics.argtypes = [v, C.c_int, i64, C.c_int32, v, v, v, v, v, v, v, v]
        lib.luf_stream_mc_metrics.argtypes = [v, C.c_int, v, u64, u64, i64, C.c_int32, v, v, v, v, v]
        lib.luf_stream_latency.argtypes = [v, C.c_int, v, u64, u64, i64, v, v]
        lib.luf_stream_latency_host.argtypes = [v, C.c_int, v, u64, u64, i64, v]
        _lib = lib
        return lib


def library():
    return _load()


def _check(rc: int):
    if rc == 0:
        return
    msg = _load().luf_last_error().decode()
    if rc == -4:
        raise EngineUnavailable(msg or "no CUDA device; the engine has no CPU fallback")
    if rc == -2:
        raise NotImplementedError(msg)
    if rc == -1:
        raise ValueError(msg)
    raise RuntimeError(f"luf_b200 error {rc}: {msg}")


def _ptr(a):
    """Raw pointer of a numpy array / torch tensor (None -> NULL)."""
    if a is None:
        return None
    if isinstance(a, np.ndarray):
        return a.ctypes.data
    return a.data_ptr()


def _stream():
    import torch
    return torch.cuda.current_stream().cuda_stream


class Engine:
    """One uploaded graph (``luf_graph``) on one CUDA device."""

    def __init__(self, code, device: int | None = None):
        lib = _load()
        if lib.luf_device_count() <= 0:
            raise EngineUnavailable("no CUDA device visible; the engine has no CPU fallback")
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', 0)) % lib.luf_device_count()
        self.device = device
        self._code = code
        self.flat = code.FLAT
        self.noise = code.NOISE
        desc, self._keep = make_desc(self.flat, NOISE_KIND[str(code.NOISE)])
        handle = C.c_void_p()
        _check(lib.luf_graph_create(C.byref(desc), device, C.byref(handle)))
        self._h = handle
        we, wd, wg = C.c_int32(), C.c_int32(), C.c_int32()
        _check(lib.luf_graph_words(self._h, C.byref(we), C.byref(wd), C.byref(wg)))
        self.WE, self.WD, self.WG = we.value, wd.value, wg.value

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib is not None:
            _lib.luf_graph_destroy(h)

    # ------------------------------------------------------------- host API
    def class_p(self, noise_level: float) -> np.ndarray:
        return np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)

    def mc_run_host(self, decoder: int, flags: int, noise_level: float, seed: int, shot0: int, nshots: int):
        """Fused sample+decode+check of ``nshots`` shots; returns (failures, shots)."""
        cp = self.class_p(noise_level)
        counts = (C.c_ulonglong * 4)(0, 0, 0, 0)
        _check(_load().luf_mc_run_host(self._h, decoder, flags, cp.ctypes.data, seed, shot0, nshots, counts))
        self.last_step_sums = (int(counts[2]), int(counts[3]))
        return int(counts[0]), int(counts[1])

    def mc_logical_host(self, decoder: int, flags: int, noise_level: float, seed: int, shot0: int, nshots: int):
        """The fused kernels with the per-shot outcome kept (``_schemes.py:143-155``): returns
        (logical uint8 [nshots], failures, shots)."""
        cp = self.class_p(noise_level)
        counts = (C.c_ulonglong * 4)(0, 0, 0, 0)
        logical = np.zeros(nshots, np.uint8)
        _check(_load().luf_mc_run_logical_host(self._h, decoder, flags, cp.ctypes.data, seed, shot0, nshots,
                                               counts, _ptr(logical)))
        self.last_step_sums = (int(counts[2]), int(counts[3]))
        return logical, int(counts[0]), int(counts[1])

    def decode_batch_host(self, decoder: int, flags: int, syn_bits, want_growth=False, want_steps=True,
                          corr_out=None):
        """Host syndromes [n, W(D)] -> host corrections [n, W(E)] (+ growth, steps); the shots stream through
        the device in chunks (copies overlap the decode when the arrays are page-locked).  ``corr_out``:
        an optional preallocated uint32 [n, W(E)] array (e.g. a view of a pinned tensor) to write into."""
        syn = np.ascontiguousarray(syn_bits, dtype=np.uint32).reshape(-1, self.WD)
        n = syn.shape[0]
        if corr_out is None:
            corr = np.zeros((n, self.WE), np.uint32)
        else:
            corr = corr_out
            if corr.dtype != np.uint32 or corr.shape != (n, self.WE) or not corr.flags.c_contiguous:
                raise ValueError(f'corr_out must be a C-contiguous uint32 array of shape {(n, self.WE)}')
        growth = np.zeros((n, self.WG), np.uint32) if want_growth else None
        steps = np.zeros((n, 2), np.int32) if want_steps else None
        _check(_load().luf_decode_batch_host(self._h, decoder, flags, n, _ptr(syn), _ptr(corr),
                                             _ptr(growth), _ptr(steps)))
        return corr, growth, steps

    def decode_frames_host(self, decoder: int, flags: int, syn_bits, want_steps=False, frame_out=None):
        """Host syndromes [n, W(D)] -> host logical frames uint8 [n] (the parity of each shot's correction on
        the logical mask; logical error = parity(error on the mask) ^ frame): one byte per shot leaves the
        device instead of W(E) words."""
        syn = np.ascontiguousarray(syn_bits, dtype=np.uint32).reshape(-1, self.WD)
        n = syn.shape[0]
        frame = np.zeros(n, np.uint8) if frame_out is None else frame_out
        if frame.dtype != np.uint8 or frame.shape != (n,) or not frame.flags.c_contiguous:
            raise ValueError(f'frame_out must be a C-contiguous uint8 array of shape {(n,)}')
        steps = np.zeros((n, 2), np.int32) if want_steps else None
        _check(_load().luf_decode_frames_host(self._h, decoder, flags, n, _ptr(syn), _ptr(frame), _ptr(steps)))
        return (frame, steps) if want_steps else frame

    def sample_host(self, noise_level: float, seed: int, shot0: int, nshots: int):
        cp = self.class_p(noise_level)
        err = np.zeros((nshots, self.WE), np.uint32)
        syn = np.zeros((nshots, self.WD), np.uint32)
        _check(_load().luf_sample_host(self._h, cp.ctypes.data, seed, shot0, nshots, _ptr(err), _ptr(syn)))
        return err, syn

    def sample_weight_host(self, weight: int, seed: int, shot0: int, nshots: int):
        """Exactly ``weight`` distinct fresh edges flipped per shot (``Noise.force_error``)."""
        err = np.zeros((nshots, self.WE), np.uint32)
        syn = np.zeros((nshots, self.WD), np.uint32)
        _check(_load().luf_sample_weight_host(self._h, int(weight), seed, shot0, nshots, _ptr(err), _ptr(syn)))
        return err, syn

    def mc_weight(self, decoder: int, flags: int, weight: int, seed: int, shot0: int, nshots: int,
                  chunk: int = 1 << 18):
        """Failures among ``nshots`` shots of fixed-weight errors: staged K1' -> K2/K3a -> K4 on
        device buffers, a chunk of shots at a time; returns (failures, shots)."""
        import torch
        dev = torch.device('cuda', self.device)
        chunk = min(chunk, max(1, nshots))
        err = torch.empty((chunk, self.WE), dtype=torch.int32, device=dev)
        syn = torch.empty((chunk, self.WD), dtype=torch.int32, device=dev)
        corr = torch.empty((chunk, self.WE), dtype=torch.int32, device=dev)
        fails = torch.zeros(1, dtype=torch.int64, device=dev)
        lib, done = _load(), 0
        with torch.cuda.device(dev):
            st = _stream()
            while done < nshots:
                k = min(chunk, nshots - done)
                _check(lib.luf_sample_weight(self._h, int(weight), seed, shot0 + done, k, _ptr(err), _ptr(syn), st))
                _check(lib.luf_decode_batch(self._h, decoder, flags, k, _ptr(syn), _ptr(corr), None, None, st))
                _check(lib.luf_check(self._h, k, _ptr(err), _ptr(corr), None, _ptr(fails), st))
                done += k
            return int(fails.item()), nshots

    def mc_global(self, decoder: int, flags: int, noise_level: float, seed: int, shot0: int, nshots: int,
                  chunk: int = 1 << 14, want_counts=False):
        """``nshots`` windows of the global batch scheme (``_schemes.py:186-203``): sample -> decode ->
        count the west-east strings of the leftover (``luf_count_strings``), all on device buffers; returns
        the total (and the per-window counts)."""
        import torch
        dev = torch.device('cuda', self.device)
        chunk = min(chunk, max(1, nshots))
        err = torch.empty((chunk, self.WE), dtype=torch.int32, device=dev)
        syn = torch.empty((chunk, self.WD), dtype=torch.int32, device=dev)
        corr = torch.empty((chunk, self.WE), dtype=torch.int32, device=dev)
        per = torch.zeros(nshots, dtype=torch.int32, device=dev)
        total = torch.zeros(1, dtype=torch.int64, device=dev)
        cp = self.class_p(noise_level)
        lib, done = _load(), 0
        with torch.cuda.device(dev):
            st = _stream()
            while done < nshots:
                k = min(chunk, nshots - done)
                _check(lib.luf_sample(self._h, cp.ctypes.data, seed, shot0 + done, k, _ptr(err), _ptr(syn), st))
                _check(lib.luf_decode_batch(self._h, decoder, flags, k, _ptr(syn), _ptr(corr), None, None, st))
                _check(lib.luf_count_strings(self._h, k, _ptr(err), _ptr(corr), per[done:].data_ptr(), _ptr(total), st))
                done += k
            out = int(total.item())
            return (out, per.cpu().numpy()) if want_counts else out

    def mc_special(self, noise_engine: 'Engine', decoder: int, flags: int, noise_level: float, seed: int, shot0: int,
                   nshots: int, chunk: int = 1 << 18):
        """``sim.accuracy.monte_carlo_special`` (``monte_carlo.py:162-201``): the errors and syndromes come from
        ``noise_engine``'s graph (circuit-level noise), this engine's graph (phenomenological, same nodes)
        decodes them; leftover parity = west parity of the error on its graph XOR west parity of the
        correction on this one.  Staged kernels on device buffers; returns (failures, shots)."""
        import torch
        a, b = noise_engine, self
        if a.WD != b.WD or a.flat.n_nodes != b.flat.n_nodes or a.flat.n_boundary != b.flat.n_boundary:
            raise ValueError('monte_carlo_special: the two codes must share their nodes')
        dev = torch.device('cuda', self.device)
        chunk = min(chunk, max(1, nshots))
        i32 = dict(dtype=torch.int32, device=dev)
        err, syn = torch.empty((chunk, a.WE), **i32), torch.empty((chunk, a.WD), **i32)
        corr = torch.empty((chunk, b.WE), **i32)
        zero_a, zero_b = torch.zeros((chunk, a.WE), **i32), torch.zeros((chunk, b.WE), **i32)
        la, lb = torch.empty(chunk, dtype=torch.uint8, device=dev), torch.empty(chunk, dtype=torch.uint8, device=dev)
        scratch = torch.zeros(1, dtype=torch.int64, device=dev)
        fails, done = 0, 0
        with torch.cuda.device(dev):
            while done < nshots:
                k = min(chunk, nshots - done)
                a.sample(noise_level, seed, shot0 + done, k, err, syn)
                b.decode_batch(decoder, flags, k, syn, corr)
                a.check(k, err, zero_a, la, scratch)
                b.check(k, zero_b, corr, lb, scratch)
                fails += int((la[:k] ^ lb[:k]).sum().item())
                done += k
        return fails, nshots

    def count_strings_host(self, err_bits, corr_bits):
        """Per-shot counts of west-east strings in ``err ^ corr`` (host arrays in, int32 [n] out)."""
        import torch
        dev = torch.device('cuda', self.device)
        e = torch.from_numpy(np.ascontiguousarray(err_bits, dtype=np.uint32).reshape(-1, self.WE).view(np.int32)).to(dev)
        c = torch.from_numpy(np.ascontiguousarray(corr_bits, dtype=np.uint32).reshape(-1, self.WE).view(np.int32)).to(dev)
        per = torch.zeros(e.shape[0], dtype=torch.int32, device=dev)
        total = torch.zeros(1, dtype=torch.int64, device=dev)
        with torch.cuda.device(dev):
            _check(_load().luf_count_strings(self._h, e.shape[0], _ptr(e), _ptr(c), _ptr(per), _ptr(total), _stream()))
            out = per.cpu().numpy()
        assert int(total.item()) == int(out.sum())
        return out

    def mc_given_errors(self, decoder: int, flags: int, err_bits, syn_bits):
        """Failures among given (error, syndrome) bit arrays: K2/K3a -> K4; returns (failures, shots)."""
        corr, _, _ = self.decode_batch_host(decoder, flags, syn_bits, want_steps=False)
        _, fails = self.check_host(err_bits, corr)
        return fails, len(corr)

    def check_host(self, err_bits, corr_bits):
        err = np.ascontiguousarray(err_bits, dtype=np.uint32).reshape(-1, self.WE)
        corr = np.ascontiguousarray(corr_bits, dtype=np.uint32).reshape(-1, self.WE)
        n = err.shape[0]
        logical = np.zeros(n, np.uint8)
        fails = C.c_ulonglong(0)
        _check(_load().luf_check_host(self._h, n, _ptr(err), _ptr(corr), _ptr(logical), C.byref(fails)))
        return logical, int(fails.value)

    # ----------------------------------------------------------- device API
    def sample(self, noise_level, seed, shot0, nshots, err_bits, syn_bits):
        cp = self.class_p(noise_level)
        _check(_load().luf_sample(self._h, cp.ctypes.data, seed, shot0, nshots,
                                  _ptr(err_bits), _ptr(syn_bits), _stream()))

    def decode_batch(self, decoder, flags, nshots, syn_bits, corr_bits, growth2b=None, steps=None):
        _check(_load().luf_decode_batch(self._h, decoder, flags, nshots, _ptr(syn_bits), _ptr(corr_bits),
                                        _ptr(growth2b), _ptr(steps), _stream()))

    def check(self, nshots, err_bits, corr_bits, logical, fail_count):
        _check(_load().luf_check(self._h, nshots, _ptr(err_bits), _ptr(corr_bits), _ptr(logical),
                                 _ptr(fail_count), _stream()))

    def mc_run(self, decoder, flags, noise_level, seed, shot0, nshots, counts):
        cp = self.class_p(noise_level)
        _check(_load().luf_mc_run(self._h, decoder, flags, cp.ctypes.data, seed, shot0, nshots,
                                  _ptr(counts), _stream()))

    @property
    def launch_count(self) -> int:
        return int(_load().luf_launch_count(self._h))

    # ------------------------------------------------------- forward scheme
    def attach_forward(self, tables=None):
        """Upload the forward-scheme bookkeeping tables (once)."""
        if getattr(self, 'forward_tables', None) is None:
            from localuf_b200 import _window
            self.forward_tables = tables if tables is not None else _window.build_forward(self._code)
            desc, self._keep_fw = make_forward_desc(self.forward_tables)
            _check(_load().luf_forward_attach(self._h, C.byref(desc)))
        return self.forward_tables

    def forward_decode_host(self, decoder: int, flags: int, init_bits, fresh_bits, want_commit=True):
        """init_bits [ns, WE], fresh_bits [ns, nc, WE] -> (steps [ns, nc, 2], commit [ns, nc, WE] | None, m [ns, nc])."""
        self.attach_forward()
        fresh = np.ascontiguousarray(fresh_bits, dtype=np.uint32)
        init = np.ascontiguousarray(init_bits, dtype=np.uint32).reshape(fresh.shape[0], self.WE)
        ns, nc = fresh.shape[0], fresh.shape[1]
        steps = np.zeros((ns, nc, 2), np.int32)
        m = np.zeros((ns, nc), np.int32)
        commit = np.zeros((ns, nc, self.WE), np.uint32) if want_commit else None
        _check(_load().luf_forward_decode_host(self._h, decoder, flags, ns, nc, _ptr(init), _ptr(fresh),
                                               _ptr(steps), _ptr(commit), _ptr(m)))
        return steps, commit, m

    def forward_mc_host(self, decoder: int, flags: int, noise_level: float, seed: int, stream0: int,
                        nstreams: int, n: int, want_steps=False):
        self.attach_forward()
        cp = self.class_p(noise_level)
        counts = (C.c_ulonglong * 4)(0, 0, 0, 0)
        steps = np.zeros((nstreams, n, 2), np.int32) if want_steps else None
        _check(_load().luf_forward_mc_host(self._h, decoder, flags, cp.ctypes.data, seed, stream0, nstreams, n,
                                           counts, _ptr(steps)))
        return [int(x) for x in counts], steps


class StreamEngine:
    """One uploaded sliding window (``luf_stream``) on one CUDA device."""

    def __init__(self, code, tables=None, device: int | None = None):
        from localuf_b200 import _window
        lib = _load()
        if lib.luf_device_count() <= 0:
            raise EngineUnavailable("no CUDA device visible; the engine has no CPU fallback")
        if device is None:
            device = int(os.environ.get('LOCAL_RANK', 0)) % lib.luf_device_count()
        self.device = device
        self.noise = code.NOISE
        self.tables = tables if tables is not None else _window.build(code)
        desc, self._keep = make_window_desc(self.tables)
        handle = C.c_void_p()
        _check(lib.luf_stream_create(C.byref(desc), device, C.byref(handle)))
        self._h = handle
        self.W = (self.tables.EL + 31) // 32

    def __del__(self):
        h, self._h = getattr(self, '_h', None), None
        if h and _lib is not None:
            _lib.luf_stream_destroy(h)

    def class_p(self, noise_level: float) -> np.ndarray:
        return np.ascontiguousarray(self.noise.class_probabilities(noise_level), dtype=np.float64)

    def decode_host(self, flags: int, fresh_bits, want_floor=True):
        """fresh_bits [nstreams, ncycles, W] -> (rt [.., 2], floor [.., W] or None, m [..])."""
        fresh = np.ascontiguousarray(fresh_bits, dtype=np.uint32)
        ns, nc = fresh.shape[0], fresh.shape[1]
        rt = np.zeros((ns, nc, 2), np.int32)
        m = np.zeros((ns, nc), np.int32)
        floor = np.zeros((ns, nc, self.W), np.uint32) if want_floor else None
        _check(_load().luf_stream_decode_host(self._h, flags, ns, nc, _ptr(fresh), _ptr(rt), _ptr(floor), _ptr(m)))
        return rt, floor, m

    def mc_host(self, flags: int, noise_level: float, seed: int, stream0: int, nstreams: int, n: int,
                want_steps=False):
        cp = self.class_p(noise_level)
        counts = (C.c_ulonglong * 4)(0, 0, 0, 0)
        steps = np.zeros((nstreams, n, 2), np.int32) if want_steps else None
        _check(_load().luf_stream_mc_host(self._h, flags, cp.ctypes.data, seed, stream0, nstreams, n, counts,
                                          _ptr(steps)))
        return [int(x) for x in counts], steps

    def latency_host(self, flags: int, noise_level: float, seed: int, stream0: int, nstreams: int):
        """``Frugal.sample_latency`` of ``nstreams`` experiments -> int32 [nstreams, 3]
        (time_only = 'all', 'merging', 'unrooting')."""
        cp = self.class_p(noise_level)
        lat = np.zeros((nstreams, 3), np.int32)
        _check(_load().luf_stream_latency_host(self._h, flags, cp.ctypes.data, seed, stream0, nstreams, _ptr(lat)))
        return lat

    # ---- per-cycle confidence-score inputs (device buffers are torch tensors: plumbing only)
    @property
    def growth_words(self) -> int:
        return int(_load().luf_stream_growth_words(self._h))

    def decode_metrics(self, flags: int, fresh_bits, drop_only=None):
        """As :meth:`decode_host`, plus the window's growth after every cycle (a DEVICE int32 tensor
        [nstreams * ncycles, growth_words], input of ``luf_dcs_run``) and mins [nstreams, ncycles, 2] =
        (min_defect_height, min_active_layer).  ``drop_only`` (optional, bool [nstreams, ncycles]): cycles on
        which the decoder only drops (``decode(defects_possible=False)``)."""
        import torch
        fresh = np.ascontiguousarray(fresh_bits, dtype=np.uint32)
        ns, nc = fresh.shape[0], fresh.shape[1]
        dev = torch.device('cuda', self.device)
        with torch.cuda.device(dev):
            fr = torch.from_numpy(fresh.view(np.int32)).to(dev)
            drop = None
            if drop_only is not None:
                drop = torch.from_numpy(np.ascontiguousarray(drop_only, dtype=np.uint8).reshape(ns, nc)).to(dev)
            rt = torch.zeros((ns, nc, 2), dtype=torch.int32, device=dev)
            m = torch.zeros((ns, nc), dtype=torch.int32, device=dev)
            floor = torch.zeros((ns, nc, self.W), dtype=torch.int32, device=dev)
            growth = torch.zeros((ns * nc, self.growth_words), dtype=torch.int32, device=dev)
            mins = torch.zeros((ns, nc, 2), dtype=torch.int32, device=dev)
            _check(_load().luf_stream_decode_metrics(self._h, flags, ns, nc, _ptr(fr), _ptr(rt), _ptr(floor), _ptr(m),
                                                     _ptr(growth), _ptr(mins), _ptr(drop), _stream()))
            return (rt.cpu().numpy(), floor.cpu().numpy().view(np.uint32), m.cpu().numpy(), growth, mins.cpu().numpy())

    def mc_metrics(self, flags: int, noise_level: float, seed: int, stream0: int, nstreams: int, n: int):
        """As :meth:`mc_host` with steps, plus the growth after every TIMED cycle (DEVICE int32 tensor
        [nstreams * n * d, growth_words]) and cycle [nstreams, n * d, 4] = (min_defect_height,
        min_active_layer, runtime 'all', runtime 'unrooting')."""
        import torch
        cp = self.class_p(noise_level)
        dev = torch.device('cuda', self.device)
        nd = n * self.tables.d
        with torch.cuda.device(dev):
            counts = torch.zeros(4, dtype=torch.int64, device=dev)
            steps = torch.zeros((nstreams, n, 2), dtype=torch.int32, device=dev)
            growth = torch.zeros((nstreams * nd, self.growth_words), dtype=torch.int32, device=dev)
            cycle = torch.zeros((nstreams, nd, 4), dtype=torch.int32, device=dev)
            _check(_load().luf_stream_mc_metrics(self._h, flags, cp.ctypes.data, seed, stream0, nstreams, n,
                                                 _ptr(counts), _ptr(steps), _ptr(growth), _ptr(cycle), _stream()))
            return [int(x) for x in counts.cpu()], steps.cpu().numpy(), growth, cycle.cpu().numpy()

    def mc(self, flags, noise_level, seed, stream0, nstreams, n, counts, steps=None):
        cp = self.class_p(noise_level)
        _check(_load().luf_stream_mc(self._h, flags, cp.ctypes.data, seed, stream0, nstreams, n,
                                     _ptr(counts), _ptr(steps), _stream()))

    @property
    def launch_count(self) -> int:
        return int(_load().luf_stream_launch_count(self._h))


def unpack_growth(growth2b: np.ndarray, n_edges: int) -> np.ndarray:
    """[n, ceil(E/16)] words of 2-bit fields -> [n, E] uint8."""
    g = np.ascontiguousarray(growth2b, dtype=np.uint32)
    b = np.unpackbits(g.view(np.uint8), axis=-1, bitorder='little')
    b = b.reshape(g.shape[0], -1, 2)
    return (b[:, :, 0] | (b[:, :, 1] << 1)).astype(np.uint8)[:, :n_edges]
