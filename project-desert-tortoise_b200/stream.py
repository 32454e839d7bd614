"""One long stream as overlapping segments (include/pdt.h `pdt_stream_plan`; SURVEY.md §8e row 2, BASELINE configs[4]).

The reference demodulates a recording strictly serially (POESTIPdemod/main.c:373-482).  Here a stream that is resident
in HBM is cut into segments which start every `segment` samples and are `lead + segment + tail` samples long; every
segment is an independent capture of one batch (`pdt_demod_segments_device`: stride = segment, per-capture lengths that
exceed the stride; segments behind the first start in track mode from a carrier estimate instead of repeating the
acquisition sweep), and the frame tables are stitched by ownership windows (`pdt_stream_stitch`).  Segments shard
across GPUs exactly like captures: rank r holds the slice of the stream its contiguous range of segments needs, and the
only exchange is the gather of the result tables before the stitch.

Everything that computes is in the C-ABI library; this module is the host-side plumbing (plan, slice ranges, gather).
"""
from __future__ import annotations

import ctypes as C
import importlib

import numpy as np

_pkg = importlib.import_module(__name__.rsplit(".", 1)[0])


class StreamPlan(C.Structure):
    _fields_ = [("total_samples", C.c_uint64), ("segment", C.c_uint64), ("lead", C.c_uint64), ("tail", C.c_uint64),
                ("n_segments", C.c_uint32), ("interp", C.c_uint32), ("seam_tol", C.c_uint32), ("pad", C.c_uint32)]


def _bind(L):
    if getattr(L, "_stream_bound", False):
        return L
    vp, u32, u64 = C.c_void_p, C.c_uint32, C.c_uint64
    L.pdt_stream_plan_make.argtypes = [C.POINTER(StreamPlan), C.POINTER(_pkg.Params), u64, u64, u64, u64]
    L.pdt_stream_segment_length.restype = u64
    L.pdt_stream_segment_length.argtypes = [C.POINTER(StreamPlan), u32]
    L.pdt_stream_stitch.restype = C.c_long
    L.pdt_stream_stitch.argtypes = [C.POINTER(StreamPlan), u32, u32, vp, vp, u32, vp, u32]
    L.pdt_demod_segments_device.argtypes = [vp, vp, C.c_int, u32, u64, vp, u32, vp]
    L.pdt_stream_frame_checks.argtypes = [vp, u32, vp]
    L.pdt_synth_poes_stream_device.argtypes = [vp, C.c_int, u64, u64, u64, C.c_double, u64, vp]
    L._stream_bound = True
    return L


def make_plan(prec: str, params, total_samples: int, segment: int, lead: int = 0, tail: int = 0) -> StreamPlan:
    L = _bind(_pkg.load(prec))
    plan = StreamPlan()
    _pkg._check(L, L.pdt_stream_plan_make(C.byref(plan), C.byref(params), total_samples, segment, lead, tail))
    return plan


def segment_lengths(prec: str, plan: StreamPlan, first: int, count: int) -> np.ndarray:
    L = _bind(_pkg.load(prec))
    return np.array([L.pdt_stream_segment_length(C.byref(plan), first + i) for i in range(count)], np.uint64)


def slice_range(plan: StreamPlan, lengths: np.ndarray, first: int) -> tuple[int, int]:
    """(start sample, sample count) of the part of the stream that segments [first, first + len(lengths)) read."""
    if len(lengths) == 0:
        return first * plan.segment, 0
    ends = (first + np.arange(len(lengths), dtype=np.uint64)) * np.uint64(plan.segment) + lengths
    start = first * plan.segment
    return int(start), int(ends.max()) - int(start)


def max_frames_per_segment(plan: StreamPlan, sample_rate: float) -> int:
    return int((plan.lead + plan.segment + plan.tail) / sample_rate * 10.0) + 8


def stitch(prec: str, plan: StreamPlan, first: int, count: int, stats: np.ndarray, frames: np.ndarray) -> np.ndarray:
    """stats[count], frames[count, max_frames] of segments [first, first+count) -> owned frames in stream order
    (FRAME_DTYPE, `sample_index` stream-global in interpolated samples)."""
    L = _bind(_pkg.load(prec))
    stats = np.ascontiguousarray(stats, _pkg.STATS_DTYPE)
    frames = np.ascontiguousarray(frames, _pkg.FRAME_DTYPE).reshape(count, -1)
    out = np.zeros(int(np.minimum(stats["n_frames"], frames.shape[1]).sum()) + 1, _pkg.FRAME_DTYPE)
    k = L.pdt_stream_stitch(C.byref(plan), first, count, _pkg._p(stats), _pkg._p(frames), frames.shape[1], _pkg._p(out), out.size)
    if k < 0:
        raise _pkg.PdtError(L.pdt_last_error().decode())
    return out[:k]


def frame_checks(prec: str, frames: np.ndarray) -> np.ndarray:
    """Parity word 103, 9-bit counter, spacecraft id and counter continuity of a stitched frame list (QUALITY_DTYPE);
    `continuous` now spans the seams between segments."""
    L = _bind(_pkg.load(prec))
    frames = np.ascontiguousarray(frames, _pkg.FRAME_DTYPE)
    q = np.zeros(frames.size, _pkg.QUALITY_DTYPE)
    _pkg._check(L, L.pdt_stream_frame_checks(_pkg._p(frames), frames.size, _pkg._p(q)))
    return q


def frame_times(frames: np.ndarray, sample_rate: float, interp: int = 1) -> np.ndarray:
    """True time of every frame's sync word in seconds (float64) from its stream-global sample index.  The reference's own
    time column (`pdt_format_frames`, which emulates it) is a float accumulated once per sample (wave.c:91-167): at 250 ksps
    it runs fast beyond 64 s and stops advancing at 128 s, so for long streams this is the column to use."""
    return frames["sample_index"].astype(np.float64) / (float(sample_rate) * max(int(interp), 1))


def continuity(frames: np.ndarray) -> dict:
    """The scale-test acceptance metric: 9-bit minor-frame counter (bytes 4-5, daytimeDecode.m:4) must
    step by one, modulo 320, from each complete frame to the next."""
    full = frames[(frames["complete"] == 1) & (frames["n_bytes"] == 104)]
    cnt = ((full["bytes"][:, 4].astype(np.int64) & 1) << 8) | full["bytes"][:, 5].astype(np.int64)
    step = (cnt[1:] - cnt[:-1]) % 320
    breaks = np.nonzero(step != 1)[0]
    return {"frames": int(frames.size), "complete": int(full.size), "counter_breaks": int(breaks.size),
            "missing_frames": int(((step[breaks] - 1) % 320).sum()) if breaks.size else 0,
            "first_break_at_frame": int(breaks[0]) + 1 if breaks.size else None}


class StreamDemod:
    """The segments [first, first + count) of a stream plan on this process's GPU."""

    def __init__(self, prec: str, params, plan: StreamPlan, first: int, count: int):
        self.prec, self.plan, self.first, self.count = prec, plan, first, count
        self.lengths = segment_lengths(prec, plan, first, count)
        self.start, self.n_slice = slice_range(plan, self.lengths, first)
        self.max_frames = max_frames_per_segment(plan, params.sample_rate)
        self.demod = _pkg.Demod(prec, params, max(count, 1), int(plan.lead + plan.segment + plan.tail), self.max_frames)

    def run_device(self, d_slice_ptr: int, pcm16: bool = False, stream: int = 0) -> None:
        """d_slice_ptr: device address of stream sample `self.start` (the slice must hold `self.n_slice` samples)."""
        if self.count:
            L = _bind(self.demod.L)
            n_serial = 1 if self.first == 0 else 0          # only the stream's first segment repeats the reference's acquisition
            _pkg._check(L, L.pdt_demod_segments_device(self.demod.ctx, d_slice_ptr, int(pcm16), self.count, int(self.plan.segment),
                                                       _pkg._p(self.lengths), n_serial, stream))

    def fetch(self, stream: int = 0):
        if not self.count:
            return np.zeros(0, _pkg.STATS_DTYPE), np.zeros((0, self.max_frames), _pkg.FRAME_DTYPE)
        return self.demod.fetch(self.count, stream)

    def stitch_local(self, stats, frames) -> np.ndarray:
        return stitch(self.prec, self.plan, self.first, self.count, stats, frames)


def gather_and_stitch(prec: str, plan: StreamPlan, stats: np.ndarray, frames: np.ndarray, counts: list[int], device=None,
                      group=None) -> np.ndarray:
    """All ranks: gather the per-rank segment tables (rank r holds counts[r] segments, in rank order) and stitch the
    whole stream.  `device` None = CPU tensors (gloo); a CUDA device = NCCL."""
    import torch
    from . import dist as pdist
    max_frames = frames.shape[1] if frames.ndim == 2 else 0
    t_s = torch.from_numpy(np.ascontiguousarray(stats).view(np.uint8).reshape(len(stats), _pkg.STATS_DTYPE.itemsize).copy())
    t_f = torch.from_numpy(np.ascontiguousarray(frames).view(np.uint8).reshape(len(stats), max_frames * _pkg.FRAME_DTYPE.itemsize).copy())
    if device is not None:
        t_s, t_f = t_s.to(device), t_f.to(device)
    g_s = pdist.gather_tables(t_s, counts, group).cpu().numpy()
    g_f = pdist.gather_tables(t_f, counts, group).cpu().numpy()
    n = sum(counts)
    all_stats = g_s.reshape(-1).view(_pkg.STATS_DTYPE)[:n]
    all_frames = g_f.reshape(-1).view(_pkg.FRAME_DTYPE).reshape(n, max_frames)
    return stitch(prec, plan, 0, n, all_stats, all_frames)
