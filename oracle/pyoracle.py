"""ctypes bindings for the CPU oracle — TEST INFRASTRUCTURE ONLY.

Two families:
  * ``Oracle(prec)``  -> oracle/liboracle_f32.so / _f64.so : our plain-C restatement (pdt_oracle.c)
  * ``RefLib(prec)``  -> oracle/_ref/libref_f32.so / _f64.so: the UNMODIFIED reference compiled from
    /root/reference (oracle/Makefile).  The reference keeps its state in function statics, so every
    ``RefLib`` instance dlopen()s a private temp copy of the .so => a fresh stream per instance.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")

POES_SYNC = b"1110110111100010000"
ARGOS_SYNC = b"0001011110000"


def build(ref: bool = True) -> None:
    """(Re)build liboracle_*.so and, when /root/reference is present, oracle/_ref/*."""
    subprocess.run(["make", "-C", HERE, "oracle"] + (["ref"] if ref else []), check=True,
                   stdout=subprocess.DEVNULL)


def _ctype(prec: str):
    return C.c_float if prec == "f32" else C.c_double


def _np(prec: str):
    return np.float32 if prec == "f32" else np.float64


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class _State:
    """Opaque zero-initialised state blob sized by pdto_sizeof()."""

    def __init__(self, lib, what: str):
        n = lib.pdto_sizeof(what.encode())
        assert n > 0, what
        self.buf = C.create_string_buffer(n)

    @property
    def p(self):
        return C.cast(self.buf, C.c_void_p)


class Oracle:
    def __init__(self, prec: str = "f32"):
        assert prec in ("f32", "f64")
        path = os.path.join(HERE, f"liboracle_{prec}.so")
        if not os.path.exists(path):
            build(ref=False)
        self.prec, self.R, self.dt = prec, _ctype(prec), _np(prec)
        self.lib = L = C.CDLL(path)
        R = self.R
        L.pdto_sizeof.restype = C.c_size_t
        L.pdto_sizeof.argtypes = [C.c_char_p]
        assert L.pdto_sizeof(b"real") == C.sizeof(R)
        L.pdto_arctan2.restype = R
        L.pdto_arctan2.argtypes = [R, R]
        L.pdto_q_rsqrt.restype = C.c_float
        L.pdto_q_rsqrt.argtypes = [C.c_float]
        L.pdto_static_gain.restype = R
        L.pdto_static_gain.argtypes = [C.c_void_p, C.c_uint, R]
        L.pdto_squelch.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, R]
        L.pdto_pll_reset.argtypes = [C.c_void_p]
        L.pdto_pll_run.restype = R
        L.pdto_pll_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint] + [R] * 6 + [C.c_void_p] * 2
        L.pdto_make_lpfir.restype = C.c_int
        L.pdto_make_lpfir.argtypes = [C.c_void_p, C.c_int, R, R, C.c_int]
        L.pdto_fir_reset.argtypes = [C.c_void_p]
        L.pdto_fir_interp_run.argtypes = [C.c_void_p] * 5 + [C.c_ulong, C.c_void_p, C.c_int, C.c_int]
        L.pdto_fir_run.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, C.c_int]
        L.pdto_agc_reset.argtypes = [C.c_void_p]
        L.pdto_agc_run.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, R, R, R, C.c_void_p]
        L.pdto_agcc_run.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, R, R]
        L.pdto_amp_run.restype = R
        L.pdto_amp_run.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, R]
        L.pdto_gardner_reset.argtypes = [C.c_void_p]
        L.pdto_gardner_run.restype = C.c_ulong
        L.pdto_gardner_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, C.c_int, R, R, R,
                                       C.c_void_p, C.c_void_p]
        L.pdto_mm_reset.argtypes = [C.c_void_p]
        L.pdto_mm_run.restype = C.c_ulong
        L.pdto_mm_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, C.c_int, R, R, R]
        L.pdto_manchester_reset.argtypes = [C.c_void_p]
        L.pdto_manchester_run.restype = C.c_ulong
        L.pdto_manchester_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, R]
        L.pdto_bytesync_reset.argtypes = [C.c_void_p]
        L.pdto_bytesync_free.argtypes = [C.c_void_p]
        L.pdto_bytesync_generic_run.restype = C.c_int
        L.pdto_bytesync_generic_run.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulong, C.c_char_p, C.c_uint, C.c_int, C.c_int]
        for f in (L.pdto_bytesync_poes_run, L.pdto_bytesync_argos_run):
            f.restype = C.c_int
            f.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_ulong, C.c_char_p, C.c_uint]
        L.pdto_chain_new.restype = C.c_void_p
        L.pdto_chain_new.argtypes = [C.c_int, C.c_double, C.c_ulong, C.c_int]
        L.pdto_chain_free.argtypes = [C.c_void_p]
        L.pdto_chain_feed.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64]
        L.pdto_chain_text.restype = C.c_void_p
        L.pdto_chain_text.argtypes = [C.c_void_p, C.POINTER(C.c_size_t)]
        L.pdto_pcm16_to_complex.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]

    # -- state factories ---------------------------------------------------------------------
    def new_state(self, what: str) -> _State:
        s = _State(self.lib, what)
        getattr(self.lib, f"pdto_{what}_reset")(s.p)
        return s

    # -- stage wrappers (numpy in / numpy out) -----------------------------------------------
    def static_gain(self, iq, desired=1.0):
        iq = np.ascontiguousarray(iq, self.dt)
        return float(self.lib.pdto_static_gain(_ptr(iq), iq.size // 2, desired))

    def pll(self, st, iq, Fs, freq_range, lock_thresh, lock_alpha, bw_acq, bw_track, want_lock=False, trace=False):
        iq = np.ascontiguousarray(iq, self.dt)
        n = iq.size // 2
        out = np.zeros(n, self.dt)
        lock = np.zeros(n, self.dt) if want_lock else None
        tp = np.zeros(n, self.dt) if trace else None
        tf = np.zeros(n, self.dt) if trace else None
        avg = self.lib.pdto_pll_run(st.p, _ptr(iq), _ptr(out), _ptr(lock), n, Fs, freq_range, lock_thresh,
                                    lock_alpha, bw_acq, bw_track, _ptr(tp), _ptr(tf))
        return out, lock, float(avg), tp, tf

    def make_lpfir(self, N, Fc, Fs, L):
        h = np.zeros(N, self.dt)
        self.lib.pdto_make_lpfir(_ptr(h), N, Fc, Fs, L)
        return h

    def fir_interp(self, st, in_time, x, h, L):
        x = np.ascontiguousarray(x, self.dt)
        n = x.size
        out = np.zeros(n * L, self.dt)
        ot = np.zeros(n * L, self.dt)
        self.lib.pdto_fir_interp_run(st.p, _ptr(in_time), _ptr(x), _ptr(out), _ptr(ot) if in_time is not None else None,
                                     n, _ptr(h), h.size, L)
        return out, ot

    def fir(self, st, x, h):
        x = np.array(x, self.dt)
        self.lib.pdto_fir_run(st.p, _ptr(x), x.size, _ptr(h), h.size)
        return x

    def agc(self, st, x, initial, attack, decay, trace=False):
        x = np.array(x, self.dt)
        tg = np.zeros(x.size, self.dt) if trace else None
        self.lib.pdto_agc_run(st.p, _ptr(x), x.size, initial, attack, decay, _ptr(tg))
        return x, tg

    def agcc(self, st, iq, initial, loop_gain):
        """NormalizingAGCC (AGC.c:164-200) on interleaved complex samples; `st` = new_state("agc")."""
        iq = np.array(iq, self.dt)
        self.lib.pdto_agcc_run(st.p, _ptr(iq), iq.size // 2, initial, loop_gain)
        return iq

    def signal_amplitude(self, avg_state, x, alpha):
        """FindSignalAmplitude (AGC.c:6-20); `avg_state` = one-element array holding the running average (in/out)."""
        x = np.ascontiguousarray(x, self.dt)
        return float(self.lib.pdto_amp_run(_ptr(avg_state), _ptr(x), x.size, alpha))

    def squelch(self, x, lock, thresh):
        x = np.array(x, self.dt)
        lock = np.ascontiguousarray(lock, self.dt)
        self.lib.pdto_squelch(_ptr(x), _ptr(lock), x.size, thresh)
        return x

    def gardner(self, st, xbuf, n, Fs, baud, step_range, kp, time=None):
        """xbuf may be longer than n (the reference reads a few stale samples past n)."""
        assert xbuf.dtype == self.dt and xbuf.size >= n + 8
        out = np.zeros(n + 8, self.dt)
        idx = np.zeros(n + 8, np.uint32)
        err = np.zeros(n + 8, self.dt)
        cnt = self.lib.pdto_gardner_run(st.p, _ptr(xbuf), _ptr(time), n, _ptr(out), Fs, baud, step_range, kp,
                                        _ptr(idx), _ptr(err))
        return out[:cnt].copy(), idx[:cnt].copy(), err[:cnt].copy()

    def mm(self, st, xbuf, n, Fs, baud, step_range, kp):
        out = np.zeros(n + 8, self.dt)
        cnt = self.lib.pdto_mm_run(st.p, _ptr(xbuf), None, n, _ptr(out), Fs, baud, step_range, kp)
        return out[:cnt].copy()

    def manchester(self, st, sym, thresh, time=None):
        sym = np.ascontiguousarray(sym, self.dt)
        bits = np.zeros(sym.size + 8, np.uint8)
        cnt = self.lib.pdto_manchester_run(st.p, _ptr(sym), _ptr(time), sym.size, _ptr(bits), thresh)
        return bits[:cnt].copy()

    def bytesync(self, st, bits, kind="poes", time=None):
        bits = np.ascontiguousarray(bits, np.uint8)
        f = self.lib.pdto_bytesync_poes_run if kind == "poes" else self.lib.pdto_bytesync_argos_run
        sync = POES_SYNC if kind == "poes" else ARGOS_SYNC
        return f(st.p, _ptr(bits), _ptr(time), bits.size, sync, len(sync))

    def bytesync_generic(self, st, bits, sync: bytes, frame_len: int, start_bit: int, time=None):
        """common/ByteSync.c:16-144 — the parameterised variant (frameLength / startBit)."""
        bits = np.ascontiguousarray(bits, np.uint8)
        t = None if time is None else np.ascontiguousarray(time, self.dt)
        return self.lib.pdto_bytesync_generic_run(st.p, _ptr(bits), _ptr(t), bits.size, sync, len(sync), frame_len, start_bit)

    def bytesync_text(self, st) -> str:
        # layout-independent accessor: text pointer/len live at the tail of pdto_bytesync; use the chain
        # accessor when available, else read via a tiny shim struct
        class BS(C.Structure):
            _fields_ = [("init", C.c_int), ("hist", C.c_char * 64), ("oldest", C.c_int), ("in_frame", C.c_int),
                        ("frame_byte_idx", C.c_int), ("bit_idx", C.c_int), ("zero", C.c_ubyte), ("one", C.c_ubyte),
                        ("byte", C.c_ubyte), ("text", C.c_void_p), ("text_len", C.c_size_t), ("text_cap", C.c_size_t)]
        bs = BS.from_buffer(st.buf)
        return C.string_at(bs.text, bs.text_len).decode() if bs.text else ""

    # -- whole chain ---------------------------------------------------------------------------
    def chain(self, iq, Fs, argos=False, chunk=None, force_min_L1=False, trace=False):
        """Run one capture through the whole chain; returns dict(text=..., totals, traces)."""
        iq = np.ascontiguousarray(iq, self.dt)
        n = iq.size // 2
        chunk = chunk or (2400 if argos else 10000)
        c = self.lib.pdto_chain_new(int(argos), float(Fs), chunk, int(force_min_L1))

        class Chain(C.Structure):
            pass
        # field offsets are not mirrored here; traces are attached through a helper table instead
        res = {}
        try:
            if trace:
                self._attach_traces(c, n, res)
            self.lib.pdto_chain_feed(c, _ptr(iq), n)
            ln = C.c_size_t(0)
            p = self.lib.pdto_chain_text(c, C.byref(ln))
            res["text"] = C.string_at(p, ln.value).decode()
            res.update(self._chain_totals(c))
            if trace:
                ns, nb = res["total_symbols"], res["total_bits"]
                for k in ("tr_sym", "tr_gerr", "tr_gidx"):
                    res[k] = res[k][:ns]
                res["tr_bits"] = res["tr_bits"][:nb]
        finally:
            self.lib.pdto_chain_free(c)
        return res

    # The chain struct is mirrored once here (kept in sync with pdt_oracle.h; a size check guards it).
    def _chain_struct(self):
        R = self.R

        class PLL(C.Structure):
            _fields_ = [("first_lock", C.c_long)] + [(k, R) for k in
                        ("damp", "alpha", "beta", "phase", "freq", "max_freq", "min_freq", "avg_phase", "locksig", "sweep")] + \
                       [("lock_freq_hz", C.c_double), ("samples_seen", C.c_uint64), ("lock_sample", C.c_uint64)]

        class FIR(C.Structure):
            _fields_ = [("init", C.c_int), ("oldest", C.c_int), ("interp_counter", C.c_ubyte), ("ring", R * 1024)]

        class AGC(C.Structure):
            _fields_ = [("init", C.c_int), ("gain", R)]

        class GAR(C.Structure):
            _fields_ = [("init", C.c_int), ("next", R), ("prev", R), ("half", R), ("step", R)]

        class MAN(C.Structure):
            _fields_ = [("clockmod", C.c_uint), ("cur", R), ("prev", R), ("prevprev", R), ("even_odd", C.c_ubyte)]

        class BS(C.Structure):
            _fields_ = [("init", C.c_int), ("hist", C.c_char * 64), ("oldest", C.c_int), ("in_frame", C.c_int),
                        ("frame_byte_idx", C.c_int), ("bit_idx", C.c_int), ("zero", C.c_ubyte), ("one", C.c_ubyte),
                        ("byte", C.c_ubyte), ("text", C.c_void_p), ("text_len", C.c_size_t), ("text_cap", C.c_size_t)]

        class CH(C.Structure):
            _fields_ = [("argos", C.c_int), ("Fs", C.c_double), ("chunk", C.c_ulong), ("L", C.c_int), ("N", C.c_int),
                        ("force_min_L1", C.c_int), ("norm_factor", R),
                        ("pll", PLL), ("fir", FIR), ("agc", AGC), ("gardner", GAR), ("man", MAN), ("sync", BS),
                        ("wave_time", R), ("wave_ts", R), ("h", C.c_void_p),
                        ("time_in", C.c_void_p), ("real_s", C.c_void_p), ("lpf", C.c_void_p), ("lpf_time", C.c_void_p),
                        ("sym", C.c_void_p), ("lock", C.c_void_p), ("bits", C.c_void_p),
                        ("chunks", C.c_uint64), ("total_samples", C.c_uint64), ("total_symbols", C.c_uint64),
                        ("total_bits", C.c_uint64), ("total_frames", C.c_uint64), ("avg_phase", R),
                        ("tr_phase", C.c_void_p), ("tr_freq", C.c_void_p), ("tr_pll_out", C.c_void_p),
                        ("tr_lpf", C.c_void_p), ("tr_agc", C.c_void_p),
                        ("tr_sym", C.c_void_p), ("tr_gerr", C.c_void_p), ("tr_gidx", C.c_void_p),
                        ("tr_bits", C.c_void_p), ("tr_sym_cap", C.c_size_t), ("tr_bits_cap", C.c_size_t)]
        assert C.sizeof(CH) == self.lib.pdto_sizeof(b"chain"), (C.sizeof(CH), self.lib.pdto_sizeof(b"chain"))
        return CH

    def _attach_traces(self, c, n, res):
        CH = self._chain_struct()
        ch = CH.from_address(c)
        L = max(ch.L, 1)
        res["tr_phase"] = np.zeros(n, self.dt)
        res["tr_freq"] = np.zeros(n, self.dt)
        res["tr_pll_out"] = np.zeros(n, self.dt)
        res["tr_lpf"] = np.zeros(n * L, self.dt)
        res["tr_agc"] = np.zeros(n * L, self.dt)
        cap = n * L // 4 + 64
        res["tr_sym"] = np.zeros(cap, self.dt)
        res["tr_gerr"] = np.zeros(cap, self.dt)
        res["tr_gidx"] = np.zeros(cap, np.uint64)
        res["tr_bits"] = np.zeros(cap, np.uint8)
        for k in ("tr_phase", "tr_freq", "tr_pll_out", "tr_lpf", "tr_agc", "tr_sym", "tr_gerr", "tr_gidx", "tr_bits"):
            setattr(ch, k, res[k].ctypes.data)
        ch.tr_sym_cap = cap
        ch.tr_bits_cap = cap

    def _chain_totals(self, c):
        ch = self._chain_struct().from_address(c)
        return dict(L=ch.L, N=ch.N, norm_factor=float(ch.norm_factor), chunks=ch.chunks,
                    total_samples=ch.total_samples, total_symbols=ch.total_symbols, total_bits=ch.total_bits,
                    total_frames=ch.total_frames, avg_phase=float(ch.avg_phase),
                    lock_sample=(ch.pll.lock_sample if ch.pll.first_lock >= 0 else -1),
                    lock_freq_hz=ch.pll.lock_freq_hz, locked=ch.pll.first_lock >= 0)

    def pcm16_to_complex(self, pcm):
        pcm = np.ascontiguousarray(pcm, np.int16)
        out = np.zeros(pcm.size, self.dt)
        self.lib.pdto_pcm16_to_complex(_ptr(pcm), pcm.size // 2, _ptr(out))
        return out


# ------------------------------------------------------------------------------------------------
def ref_available(prec: str = "f32") -> bool:
    return os.path.exists(os.path.join(REF_DIR, f"libref_{prec}.so"))


class RefLib:
    """The unmodified reference as a shared object; one private copy (= one stream) per instance."""

    def __init__(self, prec: str = "f32", fast: bool = False):
        src = os.path.join(REF_DIR, f"libref_{prec}{'_fast' if fast else ''}.so")
        if not os.path.exists(src):
            raise FileNotFoundError(src)
        self._tmpdir = tempfile.mkdtemp(prefix="pdtref_")
        self._path = os.path.join(self._tmpdir, os.path.basename(src))
        shutil.copy(src, self._path)
        self.prec, self.R, self.dt = prec, _ctype(prec), _np(prec)
        self.lib = L = C.CDLL(self._path)
        self.libc = C.CDLL(None)
        self.libc.fopen.restype = C.c_void_p
        self.libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        self.libc.fclose.argtypes = [C.c_void_p]
        R = self.R
        L.arctan2.restype = R
        L.arctan2.argtypes = [R, R]
        L.Q_rsqrt.restype = C.c_float
        L.Q_rsqrt.argtypes = [C.c_float]
        L.StaticGain.restype = R
        L.StaticGain.argtypes = [C.c_void_p, C.c_uint, R]
        L.Squelch.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, R]
        L.CarrierTrackPLL.restype = R
        L.CarrierTrackPLL.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint] + [R] * 6
        L.MakeLPFIR.restype = C.c_int
        L.MakeLPFIR.argtypes = [C.c_void_p, C.c_int, R, R, C.c_int]
        L.LowPassFilterInterp.argtypes = [C.c_void_p] * 4 + [C.c_ulong, C.c_void_p, C.c_int, C.c_int]
        L.LowPassFilter.argtypes = [C.c_void_p, C.c_ulong, C.c_void_p, C.c_int]
        L.NormalizingAGC.argtypes = [C.c_void_p, C.c_ulong, R, R, R]
        L.NormalizingAGCC.argtypes = [C.c_void_p, C.c_ulong, R, R]
        L.FindSignalAmplitude.restype = R
        L.FindSignalAmplitude.argtypes = [C.c_void_p, C.c_ulong, R]
        L.GardenerClockRecovery.restype = C.c_ulong
        L.GardenerClockRecovery.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, C.c_int, R, R, R]
        L.MMClockRecovery.restype = C.c_ulong
        L.MMClockRecovery.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, C.c_int, R, R, R]
        L.ManchesterDecode.restype = C.c_ulong
        L.ManchesterDecode.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_void_p, R]
        name = "ByteSyncOnSyncword" if prec == "f32" else "FindSyncWords"
        self._bs = getattr(L, name)
        self._bs.restype = C.c_int
        self._bs.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_char_p, C.c_uint, C.c_void_p]
        self._txt = os.path.join(self._tmpdir, "out.txt")
        self._fp = None

    def __del__(self):
        try:
            if self._fp:
                self.libc.fclose(self._fp)
            shutil.rmtree(self._tmpdir, ignore_errors=True)
        except Exception:
            pass

    def static_gain(self, iq, desired=1.0):
        iq = np.ascontiguousarray(iq, self.dt)
        return float(self.lib.StaticGain(_ptr(iq), iq.size // 2, desired))

    def pll(self, iq, Fs, freq_range, lock_thresh, lock_alpha, bw_acq, bw_track, want_lock=False):
        iq = np.ascontiguousarray(iq, self.dt)
        n = iq.size // 2
        out = np.zeros(n, self.dt)
        lock = np.zeros(n, self.dt) if want_lock else None
        # the reference printf()s " : PLL locked at …" on the latch; silence fd 1 around the call
        avg = _quiet(lambda: self.lib.CarrierTrackPLL(_ptr(iq), _ptr(out), _ptr(lock), n, Fs, freq_range, lock_thresh,
                                                      lock_alpha, bw_acq, bw_track))
        return out, lock, float(avg)

    def make_lpfir(self, N, Fc, Fs, L):
        h = np.zeros(N, self.dt)
        self.lib.MakeLPFIR(_ptr(h), N, Fc, Fs, L)
        return h

    def fir_interp(self, in_time, x, h, L):
        x = np.ascontiguousarray(x, self.dt)
        n = x.size
        out = np.zeros(n * L, self.dt)
        ot = np.zeros(n * L, self.dt)
        self.lib.LowPassFilterInterp(_ptr(in_time), _ptr(x), _ptr(out), _ptr(ot), n, _ptr(h), h.size, L)
        return out, ot

    def fir(self, x, h):
        x = np.array(x, self.dt)
        self.lib.LowPassFilter(_ptr(x), x.size, _ptr(h), h.size)
        return x

    def agc(self, x, initial, attack, decay):
        x = np.array(x, self.dt)
        self.lib.NormalizingAGC(_ptr(x), x.size, initial, attack, decay)
        return x

    def agcc(self, iq, initial, loop_gain):
        iq = np.array(iq, self.dt)
        self.lib.NormalizingAGCC(_ptr(iq), iq.size // 2, initial, loop_gain)
        return iq

    def signal_amplitude(self, x, alpha):
        x = np.ascontiguousarray(x, self.dt)
        return float(self.lib.FindSignalAmplitude(_ptr(x), x.size, alpha))

    def squelch(self, x, lock, thresh):
        x = np.array(x, self.dt)
        lock = np.ascontiguousarray(lock, self.dt)
        self.lib.Squelch(_ptr(x), _ptr(lock), x.size, thresh)
        return x

    def gardner(self, xbuf, n, Fs, baud, step_range, kp):
        """Returns (symbols, picked indices).  Indices are recovered with the SURVEY §8c trick: the time
        array carries sample indices, which the reference compacts in place."""
        assert xbuf.dtype == self.dt and xbuf.size >= n + 8
        time = np.arange(xbuf.size + 8, dtype=self.dt)
        out = np.zeros(n + 8, self.dt)
        cnt = self.lib.GardenerClockRecovery(_ptr(xbuf), _ptr(time), n, _ptr(out), Fs, baud, step_range, kp)
        return out[:cnt].copy(), time[:cnt].astype(np.uint32)

    def mm(self, xbuf, n, Fs, baud, step_range, kp):
        time = np.arange(xbuf.size + 8, dtype=self.dt)
        out = np.zeros(n + 8, self.dt)
        cnt = self.lib.MMClockRecovery(_ptr(xbuf), _ptr(time), n, _ptr(out), Fs, baud, step_range, kp)
        return out[:cnt].copy()

    def manchester(self, sym, thresh):
        sym = np.ascontiguousarray(sym, self.dt)
        time = np.zeros(sym.size + 8, self.dt)
        bits = np.zeros(sym.size + 8, np.uint8)
        cnt = self.lib.ManchesterDecode(_ptr(sym), _ptr(time), sym.size, _ptr(bits), thresh)
        return bits[:cnt].copy()

    def bytesync(self, bits, time=None):
        bits = np.ascontiguousarray(bits, np.uint8)
        if time is None:
            time = np.zeros(bits.size + 1, self.dt)
        if self._fp is None:
            self._fp = self.libc.fopen(self._txt.encode(), b"w")
        sync = POES_SYNC if self.prec == "f32" else ARGOS_SYNC
        return _quiet(lambda: self._bs(_ptr(bits), _ptr(time), bits.size, sync, len(sync), self._fp))

    def bytesync_text(self) -> str:
        if self._fp:
            self.libc.fclose(self._fp)
            self._fp = None
        with open(self._txt) as f:
            return f.read()


def _quiet(fn):
    """Run fn() with C-level stdout redirected to /dev/null (the reference printf()s progress)."""
    import sys
    sys.stdout.flush()
    libc = C.CDLL(None)
    libc.fflush(None)
    saved = os.dup(1)
    devnull = os.open(os.devnull, os.O_WRONLY)
    try:
        os.dup2(devnull, 1)
        r = fn()
        libc.fflush(None)
        return r
    finally:
        os.dup2(saved, 1)
        os.close(saved)
        os.close(devnull)


def ref_l1_available() -> bool:
    return os.path.exists(os.path.join(REF_DIR, "demodPOES_ref_L1"))


def run_ref_cli(app: str, wav_path: str, extra_args=(), exe: str | None = None):
    """Run oracle/_ref/demodPOES_ref or demodARGOS_ref (or the named binary of oracle/_ref) on a WAV file;
    returns (stdout, output-text)."""
    exe = os.path.join(REF_DIR, exe or f"demod{app}_ref")
    with tempfile.TemporaryDirectory() as d:
        p = subprocess.run([exe, *extra_args, os.path.abspath(wav_path)], cwd=d, capture_output=True, text=True,
                           errors="replace")
        outs = [f for f in os.listdir(d) if f.startswith(("minorFrames_", "packets_"))]
        text = open(os.path.join(d, outs[0])).read() if outs else ""
        return p.stdout, text


def read_wav_pcm16(path: str):
    """44-byte canonical WAV header (wave.c:303-378) -> (sample_rate, int16 array [2n])."""
    raw = open(path, "rb").read()
    rate = int.from_bytes(raw[24:28], "little")
    ch = int.from_bytes(raw[22:24], "little")
    bits = int.from_bytes(raw[34:36], "little")
    size = int.from_bytes(raw[40:44], "little")
    assert ch == 2 and bits == 16, (ch, bits)
    data = np.frombuffer(raw, np.int16, count=min(size, len(raw) - 44) // 2, offset=44)
    return rate, data[: (data.size // 2) * 2]


def ref_chain_poes(iq, fs, chunk=10000, out_path=None):
    """The reference's POES per-chunk loop (POESTIPdemod/main.c:373-482) driven through the UNMODIFIED reference
    library with preallocated buffers (bench.py's CPU arm).  Fresh static state per call.  Returns frames found; the
    minor-frame text the reference fprintf()s goes to `out_path` (default: /dev/null)."""
    r = RefLib("f32")
    lib = r.lib
    iq = np.ascontiguousarray(iq, np.float32)
    n = iq.size // 2
    Fs = np.float32(fs)
    L = int(np.rint(150000.0 / Fs))
    N = 26 * L
    if L < 1:
        return 0
    h = r.make_lpfir(N, 11000.0, np.float32(Fs * L), L)
    TWO_PI = 2.0 * np.pi
    w = TWO_PI / Fs
    time_in = (np.arange(1, chunk + 2, dtype=np.float32) / Fs).astype(np.float32)
    real_s = np.zeros(chunk, np.float32)
    lpf = np.zeros(chunk * N, np.float32)
    lpf_t = np.zeros(chunk * N, np.float32)
    sym = np.zeros(chunk * L, np.float32)
    bits = np.zeros(chunk * L, np.uint8)
    fp = r.libc.fopen((out_path or "/dev/null").encode(), b"w")
    FsL = np.float32(Fs * L)
    a_atk, a_dcy = 79.5775 * (TWO_PI / FsL), 159.1549 * (TWO_PI / FsL)
    frames = 0
    norm = 0.0
    pos = 0
    P = _ptr
    while pos < n:
        m = min(chunk, n - pos)
        x = iq[2 * pos: 2 * (pos + m)]
        if pos == 0:
            norm = lib.StaticGain(P(x), m, 1.0)
        lib.CarrierTrackPLL(P(x), P(real_s), None, m, Fs, 4500.0, 0.08, 0.3979 * w, 127.3240 * w, 10.3451 * w)
        lib.LowPassFilterInterp(P(time_in), P(real_s), P(lpf), P(lpf_t), m, P(h), N, L)
        lib.NormalizingAGC(P(lpf), m * L, norm, a_atk, a_dcy)
        ns = lib.GardenerClockRecovery(P(lpf), P(lpf_t), m * L, P(sym), int(FsL), 16640.3, 0.1, 3.0)
        nb = lib.ManchesterDecode(P(sym), P(lpf_t), ns, P(bits), 1.0)
        frames += r._bs(P(bits), P(lpf_t), nb, POES_SYNC, 19, fp)
        pos += m
    r.libc.fclose(fp)
    return frames


def ref_chain_argos(iq, fs, chunk=2400, out_path=None):
    """The reference's ARGOS per-chunk loop (ARGOSdemod/main.c:250-300) driven through the UNMODIFIED double-precision
    reference library with preallocated buffers (bench.py --mode argos CPU arm).  Fresh static state per call.  Returns
    packets found; the packet text goes to `out_path` (default /dev/null; FindSyncWords also echoes to stdout)."""
    r = RefLib("f64")
    lib = r.lib
    iq = np.ascontiguousarray(iq, np.float64)
    n = iq.size // 2
    Fs = float(fs)
    h = r.make_lpfir(50, 700.0, Fs, 1)
    w = 2.0 * np.pi / Fs
    tbuf = np.zeros(chunk + 2, np.float64)
    real_s = np.zeros(chunk, np.float64)
    lock = np.zeros(chunk, np.float64)
    sym = np.zeros(chunk, np.float64)
    bits = np.zeros(chunk, np.uint8)
    fp = r.libc.fopen((out_path or "/dev/null").encode(), b"w")
    packets, norm, pos, t = 0, 0.0, 0, 0.0
    Ts = 1.0 / Fs
    P = _ptr
    while pos < n:
        m = min(chunk, n - pos)
        x = iq[2 * pos: 2 * (pos + m)]
        tbuf[:m] = t + Ts * np.arange(1, m + 1)          # wave.c:167 accumulates time += Ts per sample (double build)
        t = float(tbuf[m - 1])
        if pos == 0:
            norm = lib.StaticGain(P(x), m, 1.0)
        lib.CarrierTrackPLL(P(x), P(real_s), P(lock), m, Fs, 550.0, 0.1, 3.1831 * w, 16.0 * w, 16.0 * w)
        lib.LowPassFilter(P(real_s), m, P(h), 50)
        lib.NormalizingAGC(P(real_s), m, norm, 79.5775 * w, 159.1549 * w)
        lib.Squelch(P(real_s), P(lock), m, 0.15)
        ns = lib.GardenerClockRecovery(P(real_s), P(tbuf), m, P(sym), int(Fs), 800.0, 0.1, 3.0)
        nb = lib.ManchesterDecode(P(sym), P(tbuf), ns, P(bits), 0.5)
        packets += r._bs(P(bits), P(tbuf), nb, ARGOS_SYNC, 13, fp)
        pos += m
    r.libc.fclose(fp)
    return packets


def ref_bytesync_generic(chunks, sync: bytes, frame_len: int, start_bit: int) -> str:
    """common/ByteSync.c:16-144 compiled alone (oracle/_ref/libref_bytesync_generic.so; neither application links it): feed
    the bit chunks [(bits uint8, time float64), …] through a private copy and return the text it wrote."""
    src = os.path.join(REF_DIR, "libref_bytesync_generic.so")
    tmp = tempfile.mkdtemp(prefix="pdtbsg_")
    try:
        path = os.path.join(tmp, "bsg.so")
        shutil.copy(src, path)
        lib = C.CDLL(path)
        libc = C.CDLL(None)
        libc.fopen.restype = C.c_void_p
        libc.fopen.argtypes = [C.c_char_p, C.c_char_p]
        libc.fclose.argtypes = [C.c_void_p]
        lib.ByteSyncOnSyncword.restype = C.c_int
        lib.ByteSyncOnSyncword.argtypes = [C.c_void_p, C.c_void_p, C.c_ulong, C.c_char_p, C.c_uint, C.c_int, C.c_int, C.c_void_p]
        out = os.path.join(tmp, "out.txt")
        fp = libc.fopen(out.encode(), b"w")
        for bits, time in chunks:
            bits = np.ascontiguousarray(bits, np.uint8)
            time = np.ascontiguousarray(time, np.float64)
            lib.ByteSyncOnSyncword(_ptr(bits), _ptr(time), bits.size, sync, len(sync), frame_len, start_bit, fp)
        libc.fclose(fp)
        return open(out).read()
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
