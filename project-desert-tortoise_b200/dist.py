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


def bind_to_gpu_numa(local_rank: int) -> dict:
    """Pin this process (and therefore the pinned host buffers it allocates next: first-touch) to the CPUs of the NUMA node
    its GPU hangs off, when the platform exposes more than one node to this process.  Returns what was found/done —
    bench.py reports it.  Pure sysfs + sched_setaffinity; a container that shows a single node is left alone."""
    import glob
    import os
    info = {"gpu_numa_node": None, "nodes_visible": 0, "bound": False}
    try:
        import torch
        pr = torch.cuda.get_device_properties(local_rank)
        bdf = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
    except Exception:
        bdf = None
    try:
        if bdf is None:
            import subprocess
            out = subprocess.run(["nvidia-smi", "--query-gpu=pci.bus_id", "--format=csv,noheader", "-i", str(local_rank)],
                                 capture_output=True, text=True, timeout=20).stdout.strip()
            bdf = out.splitlines()[0].strip() if out else None
        nodes = sorted(glob.glob("/sys/devices/system/node/node[0-9]*"))
        info["nodes_visible"] = len(nodes)
        if not bdf or len(nodes) < 2:
            return info
        bdf = bdf.lower()
        if len(bdf.split(":")[0]) == 8:            # nvidia-smi prints an 8-digit PCI domain, sysfs uses 4
            bdf = bdf[4:]
        node = int(open(f"/sys/bus/pci/devices/{bdf}/numa_node").read().strip())
        info["gpu_numa_node"] = node
        if node < 0:
            return info
        cpus = set()
        for part in open(f"/sys/devices/system/node/node{node}/cpulist").read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            info["bound"] = True
            info["cpus"] = len(allowed)
    except Exception as e:       # diagnostics only: never fail the run over topology files
        info["error"] = f"{type(e).__name__}: {e}"
    return info
