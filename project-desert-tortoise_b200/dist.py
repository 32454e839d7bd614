"""Multi-GPU plumbing of the demod path: one process per GPU, captures sharded across ranks, decoded frames gathered.

The hot path has no data-path collective: captures are independent units, so rank r simply owns a
contiguous block of the batch.  The only exchange is the gather of the fixed-stride result tables
(``pdt_capture_stats`` 96 B and ``pdt_frame`` 120 B records) at the end of a batch — NCCL over NVLink on GPUs,
gloo in the CPU tests.
"""
from __future__ import annotations

import numpy as np


def shard_range(n_items: int, rank: int, world: int) -> tuple[int, int]:
    """Contiguous block [start, start+count) of `n_items` captures owned by `rank` (sizes differ by at most one)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, base + (1 if rank < extra else 0)


def gather_tables(local: "torch.Tensor", counts: list[int], group=None) -> "torch.Tensor":
    """All-gather per-rank result tables of (possibly) different row counts into one table in capture order.

    `local` is a [rows_r, row_bytes] uint8 tensor (CPU for gloo, CUDA for NCCL); `counts[r]` is rank r's row count.
    Returns the [sum(counts), row_bytes] table on every rank."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if len(counts) != world:
        raise ValueError("counts must have one entry per rank")
    rows_max = max(counts)
    row_bytes = local.shape[1]
    pad = torch.zeros((rows_max, row_bytes), dtype=torch.uint8, device=local.device)
    pad[: local.shape[0]] = local
    out = torch.empty((world * rows_max, row_bytes), dtype=torch.uint8, device=local.device)
    dist.all_gather_into_tensor(out, pad, group=group)
    parts = [out[r * rows_max: r * rows_max + counts[r]] for r in range(world)]
    return torch.cat(parts, 0)


def frames_in_capture_order(stats: np.ndarray, frames: np.ndarray):
    """Flatten [captures, max_frames] frame tables into the list of valid frames, capture by capture."""
    out = []
    for c in range(stats.shape[0]):
        nf = min(int(stats["n_frames"][c]), frames.shape[1])
        for k in range(nf):
            out.append((c, frames[c, k]))
    return out
