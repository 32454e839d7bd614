"""pdt-b200 — B200-native IQ demodulation chain (host-side mirror of the reference interface).

The product is the pair of C-ABI shared objects built from ``csrc/`` (``libpdt_f32.so`` for the float/POES
configuration, ``libpdt_f64.so`` for the double/ARGOS configuration; include/pdt.h, include/pdt_legacy.h).
This package only loads them with ctypes and mirrors

  * the reference's stage functions, same names / argument meaning (``Legacy``: common/AGC.h, CarrierTrackPLL.h,
    LowPassFilter.h, GardenerClockRecovery.h, ManchesterDecode.h, <app>/ByteSync.h), and
  * the batch API (``Demod``) that runs whole captures through the fused sm_100a kernel.

There is no CPU fallback: loading fails loudly when the shared objects are missing, and every compute call
fails with PDT_ENODEV when no CUDA device is usable.  The package name contains a hyphen (it follows the
repository's naming); import it with ``importlib.import_module("project-desert-tortoise_b200")``.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)

PDT_MODE_POES, PDT_MODE_ARGOS = 0, 1
PDT_ENGINE_AUTO, PDT_ENGINE_EXACT, PDT_ENGINE_TILED = 0, 1, 2
PDT_CLOCK_GARDNER, PDT_CLOCK_MM = 0, 1
FRAME_MAX = 104


class PdtError(RuntimeError):
    pass


class Params(C.Structure):
    _fields_ = [("mode", C.c_int), ("sample_rate", C.c_double), ("chunk", C.c_uint32), ("interp", C.c_int),
                ("taps", C.c_int), ("force_min_interp1", C.c_int), ("max_carrier_dev", C.c_double),
                ("pll_acq_gain", C.c_double), ("pll_track_gain", C.c_double), ("pll_lock_alpha", C.c_double),
                ("pll_lock_thresh", C.c_double), ("agc_attack", C.c_double), ("agc_decay", C.c_double),
                ("lpf_fc", C.c_double), ("baud", C.c_double), ("gardner_err_lim", C.c_double),
                ("gardner_gain", C.c_double), ("manchester_resync", C.c_double), ("squelch_thresh", C.c_double),
                ("norm_factor", C.c_double), ("sync_word", C.c_char * 32), ("sync_len", C.c_int),
                ("engine", C.c_int), ("pll_warm", C.c_uint32), ("pll_tile", C.c_uint32), ("agc_min_tile", C.c_uint32),
                ("acq_first", C.c_uint32), ("sync_generic", C.c_int), ("sync_frame_len", C.c_int), ("sync_start_bit", C.c_int),
                ("clock_recovery", C.c_int), ("mm_step_range", C.c_double), ("mm_gain", C.c_double)]


class Frame(C.Structure):
    _fields_ = [("sample_index", C.c_uint64), ("bit_index", C.c_uint32), ("inverse", C.c_uint8),
                ("n_bytes", C.c_uint8), ("complete", C.c_uint8), ("pad", C.c_uint8), ("bytes", C.c_uint8 * FRAME_MAX)]


class Stats(C.Structure):
    _fields_ = [("n_samples", C.c_uint64), ("n_symbols", C.c_uint64), ("n_bits", C.c_uint64), ("n_frames", C.c_uint32),
                ("locked", C.c_int32), ("lock_sample", C.c_uint64), ("lock_freq_hz", C.c_double),
                ("norm_factor", C.c_double), ("avg_phase", C.c_double), ("final_phase", C.c_double),
                ("final_freq", C.c_double), ("final_gain", C.c_double), ("final_next", C.c_double),
                ("prelocked", C.c_int32), ("prelock_snr", C.c_float)]


class Traces(C.Structure):
    _fields_ = [("pll_phase", C.c_void_p), ("pll_freq", C.c_void_p), ("pll_out", C.c_void_p), ("lock", C.c_void_p),
                ("lpf", C.c_void_p), ("agc", C.c_void_p), ("sym", C.c_void_p), ("gardner_err", C.c_void_p),
                ("gardner_idx", C.c_void_p), ("bits", C.c_void_p), ("cap", C.c_uint64)]


FRAME_DTYPE = np.dtype([("sample_index", "<u8"), ("bit_index", "<u4"), ("inverse", "u1"), ("n_bytes", "u1"),
                        ("complete", "u1"), ("pad", "u1"), ("bytes", "u1", (FRAME_MAX,))])
assert FRAME_DTYPE.itemsize == C.sizeof(Frame) == 120
QUALITY_DTYPE = np.dtype([("counter", "<u2"), ("spacecraft", "u1"), ("parity_ok", "u1"), ("parity_bits", "u1"),
                          ("continuous", "u1"), ("valid", "u1"), ("pad", "u1")])
assert QUALITY_DTYPE.itemsize == 8
STATS_DTYPE = np.dtype([("n_samples", "<u8"), ("n_symbols", "<u8"), ("n_bits", "<u8"), ("n_frames", "<u4"),
                        ("locked", "<i4"), ("lock_sample", "<u8"), ("lock_freq_hz", "<f8"), ("norm_factor", "<f8"),
                        ("avg_phase", "<f8"), ("final_phase", "<f8"), ("final_freq", "<f8"), ("final_gain", "<f8"),
                        ("final_next", "<f8"), ("prelocked", "<i4"), ("prelock_snr", "<f4")])
assert STATS_DTYPE.itemsize == C.sizeof(Stats)


def lib_path(prec: str) -> str:
    variant = os.environ.get("PDT_LIB_VARIANT", "")          # experiment builds (tools/): libpdt_f32<variant>.so
    return os.path.join(HERE, f"libpdt_{prec}{variant if prec == 'f32' else ''}.so")


def build(verbose: bool = False) -> None:
    """Compile csrc/ for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.run(["make", "-C", os.path.join(HERE, "csrc"), "all"], check=True,
                   stdout=None if verbose else subprocess.DEVNULL)


_LIBS: dict = {}


def load(prec: str = "f32"):
    """dlopen libpdt_<prec>.so and declare every prototype of include/pdt.h and include/pdt_legacy.h."""
    if prec in _LIBS:
        return _LIBS[prec]
    path = lib_path(prec)
    if not os.path.exists(path):
        raise PdtError(f"{path} is missing - build it first (python -c 'import __graft_entry__ as g; g.build()'); "
                       "there is no CPU fallback")
    L = C.CDLL(path)
    R = C.c_float if prec == "f32" else C.c_double
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.pdt_version.restype = C.c_char_p
    L.pdt_last_error.restype = C.c_char_p
    L.pdt_real_size.restype = C.c_int
    L.pdt_device_count.restype = C.c_int
    L.pdt_set_device.argtypes = [C.c_int]
    L.pdt_params_default.argtypes = [C.POINTER(Params), C.c_int, C.c_double]
    L.pdt_create.restype = vp
    L.pdt_create.argtypes = [C.POINTER(Params), u32, u64, u32]
    L.pdt_destroy.argtypes = [vp]
    L.pdt_get_params.argtypes = [vp, C.POINTER(Params)]
    L.pdt_get_taps.argtypes = [vp, vp]
    L.pdt_demod_device.argtypes = [vp, vp, C.c_int, u32, u64, vp, vp, vp]
    L.pdt_demod_host.argtypes = [vp, vp, C.c_int, u32, u64, vp, vp, vp]
    L.pdt_demod_host_async.argtypes = [vp, vp, C.c_int, u32, u64, vp, vp]
    L.pdt_fetch.argtypes = [vp, u32, vp, vp, vp]
    L.pdt_result_tables.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(u32)]
    L.pdt_format_frames.restype = C.c_long
    L.pdt_format_frames.argtypes = [vp, vp, u32, C.c_char_p, C.c_size_t]
    L.pdt_launch_count.restype = u64
    L.pdt_set_profiling.argtypes = [vp, C.c_int]
    L.pdt_set_groups.argtypes = [vp, C.c_int]
    L.pdt_live_begin.argtypes = [vp]
    L.pdt_live_push_device.argtypes = [vp, vp, C.c_int, u32, u64, u64, vp]
    L.pdt_live_push_host.argtypes = [vp, vp, C.c_int, u32, u64, u64, vp, vp]
    L.pdt_kernel_times.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_float), C.c_int]
    L.pdt_timeline.argtypes = [vp, C.POINTER(C.c_char_p), C.POINTER(C.c_int), C.POINTER(C.c_float), C.c_int]
    L.pdt_engine.argtypes = [vp]
    L.pdt_tiled_counters.argtypes = [vp, C.POINTER(C.c_uint32 * 4), vp]
    L.pdt_synth_poes_device.argtypes = [vp, C.c_int, u32, u64, u64, C.c_double, u64, vp]
    # legacy ABI (reference names)
    L.StaticGain.restype = R
    L.StaticGain.argtypes = [vp, C.c_uint, R]
    L.FindSignalAmplitude.restype = R
    L.FindSignalAmplitude.argtypes = [vp, C.c_ulong, R]
    L.Squelch.argtypes = [vp, vp, C.c_ulong, R]
    L.NormalizingAGC.argtypes = [vp, C.c_ulong, R, R, R]
    L.NormalizingAGCC.argtypes = [vp, C.c_ulong, R, R]
    L.CarrierTrackPLL.restype = R
    L.CarrierTrackPLL.argtypes = [vp, vp, vp, C.c_uint] + [R] * 6
    L.arctan2.restype = R
    L.arctan2.argtypes = [R, R]
    L.Q_rsqrt.restype = C.c_float
    L.Q_rsqrt.argtypes = [C.c_float]
    L.sign.restype = C.c_int
    L.sign.argtypes = [R]
    L.MakeLPFIR.restype = C.c_int
    L.MakeLPFIR.argtypes = [vp, C.c_int, R, R, C.c_int]
    L.LowPassFilter.argtypes = [vp, C.c_ulong, vp, C.c_int]
    L.LowPassFilterInterp.argtypes = [vp, vp, vp, vp, C.c_ulong, vp, C.c_int, C.c_int]
    for f in (L.GardenerClockRecovery, L.MMClockRecovery):
        f.restype = C.c_ulong
        f.argtypes = [vp, vp, C.c_ulong, vp, C.c_int, R, R, R]
    L.ManchesterDecode.restype = C.c_ulong
    L.ManchesterDecode.argtypes = [vp, vp, C.c_ulong, vp, R]
    for f in (L.ByteSyncOnSyncword, L.FindSyncWords):
        f.restype = C.c_int
        f.argtypes = [vp, vp, C.c_ulong, C.c_char_p, C.c_uint, vp]
    L.pdt_legacy_reset.argtypes = []
    assert L.pdt_real_size() == C.sizeof(R)
    L._prec, L._R, L._dt = prec, R, (np.float32 if prec == "f32" else np.float64)
    _LIBS[prec] = L
    return L


EXPORTED_SYMBOLS = [
    # include/pdt.h
    "pdt_version", "pdt_last_error", "pdt_real_size", "pdt_device_count", "pdt_set_device", "pdt_params_default",
    "pdt_create", "pdt_destroy", "pdt_get_params", "pdt_get_taps", "pdt_demod_device", "pdt_demod_host", "pdt_demod_host_async",
    "pdt_fetch",
    "pdt_result_tables", "pdt_format_frames", "pdt_launch_count", "pdt_synth_poes_device", "pdt_engine",
    "pdt_demod_segments_device", "pdt_stream_plan_make", "pdt_stream_segment_length", "pdt_stream_stitch", "pdt_stream_frame_checks", "pdt_synth_poes_stream_device",
    "pdt_frame_checks", "pdt_tiled_counters", "pdt_set_profiling", "pdt_set_groups", "pdt_live_begin", "pdt_live_push_device", "pdt_live_push_host", "pdt_kernel_times", "pdt_timeline", "pdt_debug_acq_prof", "pdt_debug_chain_prof",
    # include/pdt_legacy.h
    "FindSignalAmplitude", "Squelch", "StaticGain", "NormalizingAGC", "NormalizingAGCC", "CarrierTrackPLL", "arctan2",
    "Q_rsqrt", "LowPassFilter", "LowPassFilterInterp", "MakeLPFIR", "GardenerClockRecovery", "MMClockRecovery", "sign",
    "ManchesterDecode", "ByteSyncOnSyncword", "FindSyncWords", "pdt_legacy_reset",
]


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _check(L, rc):
    if rc != 0:
        raise PdtError(f"pdt error {rc}: {L.pdt_last_error().decode()}")


def default_params(prec: str, mode: int, sample_rate: float) -> Params:
    L = load(prec)
    p = Params()
    _check(L, L.pdt_params_default(C.byref(p), mode, float(sample_rate)))
    return p


class Demod:
    """Batch context: many captures -> frames, one fused kernel per batch (include/pdt.h)."""

    def __init__(self, prec: str, params: Params, max_captures: int, max_samples: int, max_frames: int = 64):
        self.L = load(prec)
        self.prec, self.dt = prec, self.L._dt
        self.max_captures, self.max_samples, self.max_frames = max_captures, max_samples, max_frames
        self.ctx = self.L.pdt_create(C.byref(params), max_captures, max_samples, max_frames)
        if not self.ctx:
            raise PdtError(f"pdt_create failed: {self.L.pdt_last_error().decode()}")
        self.params = Params()
        _check(self.L, self.L.pdt_get_params(self.ctx, C.byref(self.params)))

    def close(self):
        if getattr(self, "ctx", None):
            self.L.pdt_destroy(self.ctx)
            self.ctx = None

    __del__ = close

    def taps(self) -> np.ndarray:
        h = np.zeros(max(self.params.taps, 0), self.dt)
        _check(self.L, self.L.pdt_get_taps(self.ctx, _p(h)))
        return h

    def demod_host(self, iq: np.ndarray, n_captures: int = 1, pcm16: bool = False, n_samples=None):
        """iq: [n_captures, stride, 2] (or flat) host array of REAL (or int16 when pcm16)."""
        iq = np.ascontiguousarray(iq, np.int16 if pcm16 else self.dt)
        stride = iq.size // (2 * n_captures)
        stats = np.zeros(n_captures, STATS_DTYPE)
        frames = np.zeros((n_captures, self.max_frames), FRAME_DTYPE)
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, np.uint64)
        _check(self.L, self.L.pdt_demod_host(self.ctx, _p(iq), int(pcm16), n_captures, stride, _p(ns), _p(stats), _p(frames)))
        return stats, frames

    def demod_host_async(self, iq: np.ndarray, n_captures: int = 1, pcm16: bool = False, n_samples=None, stream: int = 0):
        """Enqueue H2D + kernels for a host batch and return; collect with fetch(n_captures, stream).  `iq` must stay alive
        (and should be pinned) until then."""
        assert iq.flags.c_contiguous and iq.dtype == (np.int16 if pcm16 else self.dt)
        stride = iq.size // (2 * n_captures)
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, np.uint64)
        _check(self.L, self.L.pdt_demod_host_async(self.ctx, _p(iq), int(pcm16), n_captures, stride, _p(ns), C.c_void_p(stream)))

    def demod_device(self, d_ptr: int, n_captures: int, stride: int, pcm16: bool = False, n_samples=None, traces=None,
                     stream: int = 0):
        ns = None if n_samples is None else np.ascontiguousarray(n_samples, np.uint64)
        tr = None
        if traces is not None:
            tr = (Traces * n_captures)(*traces)
        _check(self.L, self.L.pdt_demod_device(self.ctx, d_ptr, int(pcm16), n_captures, stride, _p(ns),
                                               C.cast(tr, C.c_void_p) if tr is not None else None, stream))

    def fetch(self, n_captures: int, stream: int = 0, want_frames: bool = True):
        stats = np.zeros(n_captures, STATS_DTYPE)
        frames = np.zeros((n_captures, self.max_frames), FRAME_DTYPE) if want_frames else None
        _check(self.L, self.L.pdt_fetch(self.ctx, n_captures, _p(stats), _p(frames), stream))
        return stats, frames

    @property
    def engine(self) -> int:
        return int(self.L.pdt_engine(self.ctx))

    def tiled_counters(self, stream: int = 0):
        """(PLL tiles re-run, AGC tiles re-run, acquisition restarts, max PLL tiles per capture) of the last batch."""
        out = (C.c_uint32 * 4)()
        _check(self.L, self.L.pdt_tiled_counters(self.ctx, C.byref(out), stream))
        return tuple(int(v) for v in out)

    def set_groups(self, max_groups: int):
        _check(self.L, self.L.pdt_set_groups(self.ctx, int(max_groups)))

    def set_profiling(self, on: bool = True):
        _check(self.L, self.L.pdt_set_profiling(self.ctx, int(on)))

    def kernel_times(self):
        """[(kernel name, ms)] of the last batch (tiled engine with profiling enabled), in launch order."""
        names = (C.c_char_p * 64)()
        ms = (C.c_float * 64)()
        k = self.L.pdt_kernel_times(self.ctx, names, ms, 64)
        if k < 0:
            raise PdtError(self.L.pdt_last_error().decode())
        return [(names[i].decode(), float(ms[i])) for i in range(k)]

    def timeline(self):
        """[(kernel name, stream group, end ms)] of the last batch run with set_profiling(2)."""
        cap = 512
        names = (C.c_char_p * cap)()
        groups = (C.c_int * cap)()
        ms = (C.c_float * cap)()
        k = self.L.pdt_timeline(self.ctx, names, groups, ms, cap)
        if k < 0:
            raise PdtError(self.L.pdt_last_error().decode())
        return [(names[i].decode(), int(groups[i]), float(ms[i])) for i in range(k)]

    def frame_checks(self, n_captures: int, stream: int = 0) -> np.ndarray:
        """[n_captures, max_frames] quality records of the last batch (parity word 103, counter continuity, spacecraft id)."""
        q = np.zeros((n_captures, self.max_frames), QUALITY_DTYPE)
        _check(self.L, self.L.pdt_frame_checks(self.ctx, n_captures, _p(q), C.c_void_p(stream)))
        return q

    def result_tables(self):
        ds, df, mf = C.c_void_p(), C.c_void_p(), C.c_uint32()
        _check(self.L, self.L.pdt_result_tables(self.ctx, C.byref(ds), C.byref(df), C.byref(mf)))
        return ds.value, df.value, mf.value

    def format_frames(self, frames_1d: np.ndarray, n_frames: int) -> str:
        fr = np.ascontiguousarray(frames_1d[:n_frames])
        buf = C.create_string_buffer(400 * max(n_frames, 1) + 64)
        n = self.L.pdt_format_frames(self.ctx, _p(fr), n_frames, buf, len(buf))
        if n < 0:
            raise PdtError(self.L.pdt_last_error().decode())
        return buf.raw[:n].decode()


def frames_to_text_rows(frames_1d: np.ndarray, n_frames: int):
    """[(inverse, bytes ndarray)] for the first n_frames slots of one capture."""
    rows = []
    for f in frames_1d[:n_frames]:
        rows.append((bool(f["inverse"]), np.array(f["bytes"][: f["n_bytes"]], np.uint8), bool(f["complete"])))
    return rows


class Live:
    """Bounded-latency streaming (include/pdt.h pdt_live_*): the reference's sound-card loop
    (POESTIPdemodPortAudio/main.c:324-401) for `n_streams` streams at once.  push() hands the next chunk of every stream to the
    chain and returns, per stream, the frames COMPLETED since the previous push, in order."""

    def __init__(self, prec: str, params: Params, n_streams: int, max_chunk: int, max_frames: int = 64):
        self.d = Demod(prec, params, n_streams, max_chunk, max_frames)
        self.n_streams, self.max_chunk, self.max_frames = n_streams, max_chunk, max_frames
        _check(self.d.L, self.d.L.pdt_live_begin(self.d.ctx))
        self.reported = [0] * n_streams
        self.stats = None

    def push(self, iq: np.ndarray, pcm16: bool = False):
        """iq: [n_streams, n, 2] (or [n_streams, 2n]) — the next n <= max_chunk samples of every stream."""
        iq = np.ascontiguousarray(iq, np.int16 if pcm16 else self.d.dt).reshape(self.n_streams, -1)
        n = iq.shape[1] // 2
        assert n <= self.max_chunk
        stats = np.zeros(self.n_streams, STATS_DTYPE)
        frames = np.zeros((self.n_streams, self.max_frames), FRAME_DTYPE)
        _check(self.d.L, self.d.L.pdt_live_push_host(self.d.ctx, _p(iq), int(pcm16), self.n_streams, n, n, _p(stats), _p(frames)))
        self.stats = stats
        out = []
        for s in range(self.n_streams):
            done = []
            while self.reported[s] < int(stats[s]["n_frames"]):
                f = frames[s, self.reported[s] % self.max_frames]
                if not f["complete"]:
                    break                                   # still being shifted in: a later push finishes it
                done.append(f.copy())
                self.reported[s] += 1
            out.append(done)
        return out

    def pending(self):
        """Frames started but not complete (e.g. at the end of the input), per stream."""
        return [int(self.stats[s]["n_frames"]) - self.reported[s] if self.stats is not None else 0 for s in range(self.n_streams)]

    def close(self):
        self.d.close()


class Legacy:
    """The reference's stage functions (same names, argument meaning and in-place behaviour), numpy in/out.

    One process-wide stream, like the reference; ``reset()`` is the only addition.
    """

    def __init__(self, prec: str = "f32"):
        self.L = load(prec)
        self.prec, self.dt, self.R = prec, self.L._dt, self.L._R
        self.libc = C.CDLL(None)
        self.libc.fopen.restype = C.c_void_p
        self.libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        self.libc.fclose.argtypes = [C.c_void_p]

    def reset(self):
        self.L.pdt_legacy_reset()

    def StaticGain(self, iq, desired=1.0):
        iq = np.ascontiguousarray(iq, self.dt)
        return float(self.L.StaticGain(_p(iq), iq.size // 2, desired))

    def CarrierTrackPLL(self, iq, Fs, freq_range, lock_thresh, lock_alpha, bw_acq, bw_track, want_lock=False):
        iq = np.ascontiguousarray(iq, self.dt)
        n = iq.size // 2
        out = np.zeros(n, self.dt)
        lock = np.zeros(n, self.dt) if want_lock else None
        avg = self.L.CarrierTrackPLL(_p(iq), _p(out), _p(lock), n, Fs, freq_range, lock_thresh, lock_alpha, bw_acq, bw_track)
        return out, lock, float(avg)

    def MakeLPFIR(self, N, Fc, Fs, L):
        h = np.zeros(N, self.dt)
        self.L.MakeLPFIR(_p(h), N, Fc, Fs, L)
        return h

    def LowPassFilterInterp(self, in_time, x, h, L):
        x = np.ascontiguousarray(x, self.dt)
        n = x.size
        out = np.zeros(n * L, self.dt)
        ot = np.zeros(n * L, self.dt)
        self.L.LowPassFilterInterp(_p(in_time), _p(x), _p(out), _p(ot), n, _p(h), h.size, L)
        return out, ot

    def LowPassFilter(self, x, h):
        x = np.array(x, self.dt)
        self.L.LowPassFilter(_p(x), x.size, _p(h), h.size)
        return x

    def NormalizingAGC(self, x, initial, attack, decay):
        x = np.array(x, self.dt)
        self.L.NormalizingAGC(_p(x), x.size, initial, attack, decay)
        return x

    def NormalizingAGCC(self, iq, initial, loop_gain):
        """AGC.c:164-200 on interleaved complex samples (returned copy is the in-place result)."""
        iq = np.array(iq, self.dt)
        self.L.NormalizingAGCC(_p(iq), iq.size // 2, initial, loop_gain)
        return iq

    def FindSignalAmplitude(self, x, alpha):
        """AGC.c:6-20: running average of |x| carried across calls."""
        x = np.ascontiguousarray(x, self.dt)
        return float(self.L.FindSignalAmplitude(_p(x), x.size, alpha))

    def Squelch(self, x, lock, thresh):
        x = np.array(x, self.dt)
        lock = np.ascontiguousarray(lock, self.dt)
        self.L.Squelch(_p(x), _p(lock), x.size, thresh)
        return x

    def GardenerClockRecovery(self, xbuf, n, Fs, baud, step_range, kp, mm=False):
        assert xbuf.dtype == self.dt and xbuf.size >= n + 16
        time = np.arange(xbuf.size + 8, dtype=self.dt)
        out = np.zeros(n + 8, self.dt)
        f = self.L.MMClockRecovery if mm else self.L.GardenerClockRecovery
        cnt = f(_p(xbuf), _p(time), n, _p(out), Fs, baud, step_range, kp)
        return out[:cnt].copy(), time[:cnt].astype(np.uint32)

    def ManchesterDecode(self, sym, thresh):
        sym = np.ascontiguousarray(sym, self.dt)
        time = np.zeros(sym.size + 8, self.dt)
        bits = np.zeros(sym.size + 8, np.uint8)
        cnt = self.L.ManchesterDecode(_p(sym), _p(time), sym.size, _p(bits), thresh)
        return bits[:cnt].copy()

    def ByteSync(self, bits, path, time=None, argos=False, sync=None):
        bits = np.ascontiguousarray(bits, np.uint8)
        if time is None:
            time = np.zeros(bits.size + 1, self.dt)
        fp = self.libc.fopen(path.encode(), b"a")
        try:
            fn = self.L.FindSyncWords if argos else self.L.ByteSyncOnSyncword
            sync = sync or (b"0001011110000" if argos else b"1110110111100010000")
            return fn(_p(bits), _p(time), bits.size, sync, len(sync), fp)
        finally:
            self.libc.fclose(fp)
