"""Multi-process (world_size 2 and 3, gloo, CPU) tests of the batch sharding used by the multi-GPU path:
contiguous blocks by comparison index, no data-path collective, one final all-gather of result rows."""
import os
import socket

import numpy as np
import pytest
import torch.multiprocessing as mp

from protocols.secure_comparison_b200.sharding import all_gather_rows, shard_arrays, shard_range, shard_sizes


def test_shard_ranges_partition_the_batch():
    for total in (0, 1, 7, 64, 65536, 1048576 + 3):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_range(total, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = shard_sizes(total, world)
            assert sum(sizes) == total and max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 2, 2)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, total, ret):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(5)
        work = {"x": rng.integers(0, 1 << 32, size=(total, 4), dtype=np.uint64).astype(np.uint32),
                "flag": rng.integers(0, 2, size=(total,), dtype=np.uint8)}
        mine = shard_arrays(work, rank, world)
        lo, hi = shard_range(total, rank, world)
        assert mine["x"].shape[0] == hi - lo
        # stand-in for the per-rank device pass: an index-dependent transformation of the local rows only
        local = mine["x"] ^ np.arange(lo, hi, dtype=np.uint32)[:, None]
        full = all_gather_rows(local, total)
        want = work["x"] ^ np.arange(total, dtype=np.uint32)[:, None]
        flags = all_gather_rows(mine["flag"], total)
        ret[rank] = bool(np.array_equal(full, want) and np.array_equal(flags, work["flag"]))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("world,total", [(2, 64), (2, 33), (3, 10)])
def test_all_gather_rows_over_gloo(world, total):
    ctx = mp.get_context("spawn")
    with ctx.Manager() as mgr:
        ret = mgr.dict()
        port = _free_port()
        procs = [ctx.Process(target=_worker, args=(r, world, port, total, ret)) for r in range(world)]
        for p in procs:
            p.start()
        for p in procs:
            p.join(120)
        assert all(p.exitcode == 0 for p in procs)
        assert dict(ret) == {r: True for r in range(world)}
