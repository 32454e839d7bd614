"""Host logic of the stream mode (include/pdt.h pdt_stream_*): segment geometry, ownership windows, stitch, and the
world_size-2 gloo gather.  No device work: pdt_stream_plan_make / _segment_length / _stitch are host-only entry points."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pdt = importlib.import_module("project-desert-tortoise_b200")
stream = importlib.import_module("project-desert-tortoise_b200.stream")


def _params(fs=250000):
    return pdt.default_params("f32", pdt.PDT_MODE_POES, fs)


@pytest.mark.parametrize("total,segment,lead,tail", [(10_000_000, 1_000_000, 0, 0), (1_000_000, 1_000_000, 0, 0), (100, 1000, 0, 0),
                                                     (3_333_333, 500_000, 100_000, 40_000), (2_150_000, 1_000_000, 150_000, 36_596)])
def test_plan_geometry(total, segment, lead, tail):
    p = _params()
    plan = stream.make_plan("f32", p, total, segment, lead, tail)
    assert plan.lead == (lead or 75000) and plan.tail == (tail or 32500 + 4096) and plan.interp == 1
    k = plan.n_segments
    lens = stream.segment_lengths("f32", plan, 0, k)
    starts = np.arange(k, dtype=np.uint64) * np.uint64(segment)
    assert k >= 1 and np.all(lens >= 1) and np.all(starts + lens <= total)
    assert int(starts[-1] + lens[-1]) == total                                 # the last segment runs to the end of the stream
    # ownership windows tile [0, total): segment s owns [s·seg + lead, (s+1)·seg + lead), first from 0, last to the end,
    # and every window but the last is followed by at least `tail` samples inside its segment
    for s in range(k - 1):
        hi = (s + 1) * segment + plan.lead
        assert hi + plan.tail <= int(starts[s] + lens[s])
        assert hi < total
    if k > 1:
        assert (k - 1) * segment + plan.lead < total                           # the last window is not empty
    s0, n0 = stream.slice_range(plan, lens[1:], 1) if k > 1 else (0, 0)
    if k > 1:
        assert s0 == segment and s0 + n0 == total


def _fake_tables(plan, rng, max_frames=64):
    """Every segment "decodes" a frame every 25 000 interpolated samples of the stream (same global grid for all segments),
    starting somewhere inside its lead; bytes carry the global frame number."""
    k, L, period = plan.n_segments, plan.interp, 25000
    stats = np.zeros(k, pdt.STATS_DTYPE)
    frames = np.zeros((k, max_frames), pdt.FRAME_DTYPE)
    lens = stream.segment_lengths("f32", plan, 0, k)
    for s in range(k):
        start = s * plan.segment * L
        end = start + int(lens[s]) * L
        first = start + (0 if s == 0 else int(rng.integers(0, plan.lead * L // 2)))
        g = -(-first // period) * period + 7
        n = 0
        while g + period <= end and n < max_frames:                              # only frames that complete inside the segment
            f = frames[s, n]
            f["sample_index"], f["n_bytes"], f["complete"] = g - start, 104, 1
            num = g // period
            f["bytes"][4], f["bytes"][5] = (num % 320) >> 8, (num % 320) & 0xFF
            f["bytes"][6:14] = np.frombuffer(np.uint64(num).tobytes(), np.uint8)
            g += period
            n += 1
        stats[s]["n_frames"] = n
    return stats, frames


def test_stitch_keeps_every_frame_exactly_once():
    rng = np.random.default_rng(1)
    plan = stream.make_plan("f32", _params(), 10_000_000, 1_000_000)
    stats, frames = _fake_tables(plan, rng)
    out = stream.stitch("f32", plan, 0, plan.n_segments, stats, frames)
    nums = np.array([int(np.frombuffer(f["bytes"][6:14].tobytes(), np.uint64)[0]) for f in out])
    assert np.array_equal(nums, np.arange(nums[0], nums[0] + nums.size))          # no duplicate, no hole, stream order
    assert nums[0] == 0 and nums[-1] == (10_000_000 - 7) // 25000 - 1
    assert np.array_equal(out["sample_index"], nums * 25000 + 7)                  # stream-global positions
    c = stream.continuity(out)
    assert c["counter_breaks"] == 0 and c["complete"] == nums.size
    q = stream.frame_checks("f32", out)                                         # host twin of the device post-checks
    assert q["valid"].all() and q["continuous"].all() and np.array_equal(q["counter"], nums % 320)
    from tests.synth_ref import check_parity
    assert np.array_equal(q["parity_ok"].astype(bool), np.array([check_parity(f["bytes"]) for f in out]))
    assert 0 < q["parity_ok"].sum() < q.size                                    # random fake payload: both outcomes occur
    t = stream.frame_times(out, 250000.0, plan.interp)
    assert np.allclose(np.diff(t), 0.1) and abs(t[0] - 7 / 250000.0) < 1e-12
    # a segment that failed to lock inside its lead shows up as a counter break, not as silent loss
    stats2 = stats.copy()
    stats2[4]["n_frames"] = 0
    out2 = stream.stitch("f32", plan, 0, plan.n_segments, stats2, frames)
    c2 = stream.continuity(out2)
    # (39 of segment 4's 40 frames: its first one lies 7 samples behind the edge, inside the seam tolerance, and is kept
    # from segment 3's copy)
    assert c2["counter_breaks"] == 1 and c2["missing_frames"] == 39
    assert int((stream.frame_checks("f32", out2)["continuous"] == 0).sum()) == 1
    # stitching ranges separately and concatenating is the same as stitching everything (what the ranks do)
    a = stream.stitch("f32", plan, 0, 4, stats[:4], frames[:4])
    b = stream.stitch("f32", plan, 4, plan.n_segments - 4, stats[4:], frames[4:])
    assert np.array_equal(np.concatenate([a, b]), out)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pd = importlib.import_module("project-desert-tortoise_b200.dist")
    st = importlib.import_module("project-desert-tortoise_b200.stream")
    pk = importlib.import_module("project-desert-tortoise_b200")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        plan = st.make_plan("f32", pk.default_params("f32", pk.PDT_MODE_POES, 250000), 7_300_000, 1_000_000)
        stats, frames = _fake_tables(plan, np.random.default_rng(5))               # same seed: every rank knows the "truth"
        first, cnt = pd.shard_range(plan.n_segments, rank, world)
        counts = [pd.shard_range(plan.n_segments, r, world)[1] for r in range(world)]
        out = st.gather_and_stitch("f32", plan, stats[first:first + cnt], frames[first:first + cnt], counts)
        want = st.stitch("f32", plan, 0, plan.n_segments, stats, frames)
        q.put((rank, first, cnt, bool(np.array_equal(out, want)), int(out.size)))
    finally:
        dist.destroy_process_group()


def test_stream_shards_and_stitch_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res[0][1] == 0 and res[0][1] + res[0][2] == res[1][1]                   # contiguous segment ranges
    assert all(r[3] for r in res) and res[0][4] == res[1][4] > 250


def test_plan_edge_cases_and_stitch_capacity():
    p = _params()
    # a stream shorter than the lead is one segment that owns everything
    plan = stream.make_plan("f32", p, 50_000, 1_000_000)
    assert plan.n_segments == 1 and int(stream.segment_lengths("f32", plan, 0, 1)[0]) == 50_000
    # window of the last segment may be tiny but is never empty, and no segment starts at or behind the end of the stream
    for total in (1_075_001, 1_075_000, 2_000_000, 2_074_999):
        plan = stream.make_plan("f32", p, total, 1_000_000)
        k = plan.n_segments
        assert (k - 1) * plan.segment + (plan.lead if k > 1 else 0) < total
        assert k * plan.segment + plan.lead >= total
    # bad arguments are refused by the library, not silently planned
    with pytest.raises(pdt.PdtError):
        stream.make_plan("f32", p, 0, 1000)
    with pytest.raises(pdt.PdtError):
        stream.make_plan("f32", p, 1000, 0)
    # the stitched table must fit: too small an output is an error, not a truncation
    import ctypes as C
    plan = stream.make_plan("f32", p, 3_000_000, 1_000_000)
    stats, frames = _fake_tables(plan, np.random.default_rng(2))
    L = stream._bind(pdt.load("f32"))
    out = np.zeros(5, pdt.FRAME_DTYPE)
    rc = L.pdt_stream_stitch(C.byref(plan), 0, plan.n_segments, pdt._p(stats), pdt._p(frames), frames.shape[1], pdt._p(out), out.size)
    assert rc < 0 and b"too small" in L.pdt_last_error()
    # a range that runs past the plan is refused
    assert L.pdt_stream_stitch(C.byref(plan), 1, plan.n_segments, pdt._p(stats), pdt._p(frames), frames.shape[1], pdt._p(out), out.size) < 0


@pytest.mark.parametrize("jit_a,jit_b", [(0, 0), (-1, 0), (0, -1), (1, 0), (0, 1), (-9, 9), (9, -9), (-1, 1), (15, -14)])
def test_stitch_frame_exactly_on_a_seam_with_position_jitter(jit_a, jit_b):
    """neighbouring segments see the same sync word up to a symbol apart (own chunk grid, chunk-relative float
    Gardner state).  A frame whose sync lands exactly on an ownership edge, seen by segment s at edge + jit_a and by segment
    s+1 at edge + jit_b, must come out exactly once whatever the signs — never twice, never lost."""
    plan = stream.make_plan("f32", _params(), 4_000_000, 1_000_000)
    assert plan.seam_tol == int(np.ceil(2 * 250000 / 16640.3))                  # two symbols
    k, period = plan.n_segments, 25000
    edge = 2 * plan.segment + plan.lead                                         # ownership edge between segments 1 and 2
    stats = np.zeros(k, pdt.STATS_DTYPE)
    frames = np.zeros((k, 64), pdt.FRAME_DTYPE)
    lens = stream.segment_lengths("f32", plan, 0, k)
    for s in range(k):
        start, end = s * plan.segment, s * plan.segment + int(lens[s])
        n = 0
        for num in range(-200, 400):
            g = edge + num * period                                             # the global frame grid passes through the edge
            if g < start + (0 if s == 0 else 1000) or g + period > end:
                continue
            jit = jit_a if s == 1 else (jit_b if s == 2 else 0)
            f = frames[s, n]
            f["sample_index"], f["n_bytes"], f["complete"] = g + (jit if num == 0 else 0) - start, 104, 1
            f["bytes"][6:14] = np.frombuffer(np.int64(num).tobytes(), np.uint8)
            n += 1
        stats[s]["n_frames"] = n
    out = stream.stitch("f32", plan, 0, k, stats, frames)
    nums = np.array([int(np.frombuffer(f["bytes"][6:14].tobytes(), np.int64)[0]) for f in out])
    assert np.array_equal(nums, np.arange(nums[0], nums[0] + nums.size)), nums   # every frame once, in stream order
    assert nums[0] < 0 < nums[-1]
    at = int(np.nonzero(nums == 0)[0][0])
    assert int(out[at]["sample_index"]) in (edge + jit_a, edge + jit_b)
