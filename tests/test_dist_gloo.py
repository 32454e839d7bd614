"""N>1 host logic on CPU: world_size-2 gloo run of the capture sharding + result-table gather used by bench.py."""
import importlib
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n_caps, row_bytes, q):
    import torch
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    pd = importlib.import_module("project-desert-tortoise_b200.dist")
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        start, cnt = pd.shard_range(n_caps, rank, world)
        # the "decoded" table of this rank: row i of capture c carries c in every byte position pattern
        local = torch.empty((cnt, row_bytes), dtype=torch.uint8)
        for i in range(cnt):
            local[i] = torch.arange(row_bytes, dtype=torch.int64).add(start + i).remainder(251).to(torch.uint8)
        counts = [pd.shard_range(n_caps, r, world)[1] for r in range(world)]
        table = pd.gather_tables(local, counts)
        q.put((rank, start, cnt, table.numpy().copy()))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_caps", [7, 8, 1])
def test_shard_and_gather_world2(n_caps):
    import torch.multiprocessing as mp
    world, row_bytes = 2, 120
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n_caps, row_bytes, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = np.stack([(np.arange(row_bytes) + c) % 251 for c in range(n_caps)]).astype(np.uint8)
    covered = []
    for rank, start, cnt, table in sorted(res):
        covered += list(range(start, start + cnt))
        assert np.array_equal(table, want), rank          # every rank holds the whole table in capture order
    assert covered == list(range(n_caps))                 # shards tile the batch exactly once


def test_shard_range_properties():
    pd = importlib.import_module("project-desert-tortoise_b200.dist")
    for n in (0, 1, 5, 1024, 1025):
        for w in (1, 2, 3, 8):
            spans = [pd.shard_range(n, r, w) for r in range(w)]
            assert sum(c for _, c in spans) == n
            assert all(spans[i][0] + spans[i][1] == spans[i + 1][0] for i in range(w - 1))
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1
    with pytest.raises(ValueError):
        pd.shard_range(4, 2, 2)
