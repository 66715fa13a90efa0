"""CPU, world_size 2, gloo: the host-side data-parallel logic (row sharding + flat-bucket mean all-reduce)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from safe_dreamer_b200.parallel import GradBucket, shard_rows


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    shapes = {"a.weight": (4, 3), "a.bias": (4,), "blk.weight": (2, 5, 2)}
    b = GradBucket(shapes, "cpu")
    lo, hi = shard_rows(10, rank, world)
    rows = torch.arange(10, dtype=torch.float32)[lo:hi]
    for n in b.names:                       # "local gradient" = sum over this rank's rows
        b.views[n] += rows.sum() * (1 + b.names.index(n))
    b.allreduce_async()
    b.wait()
    out[rank] = {n: b.views[n].clone().numpy() for n in b.names}
    dist.destroy_process_group()


def test_shard_rows_partition():
    for rows in (1, 7, 16, 128):
        for world in (1, 2, 3, 8):
            spans = [shard_rows(rows, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == rows
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_bucket_views_alias_flat():
    b = GradBucket({"x": (2, 3), "y": (5,)}, "cpu")
    b.views["x"].fill_(2.0)
    b.views["y"].fill_(3.0)
    assert b.flat.tolist() == [2.0] * 6 + [3.0] * 5
    b.zero_()
    assert float(b.views["y"].abs().sum()) == 0.0
    b.allreduce_async()  # no process group: no-op
    b.wait()


def test_allreduce_mean_world2():
    world, port = 2, _free_port()
    with mp.Manager() as m:
        out = m.dict()
        mp.spawn(_worker, args=(world, port, out), nprocs=world, join=True)
        res = dict(out)
    total = float(np.arange(10).sum())
    for rank in range(world):
        for i, n in enumerate(["a.weight", "a.bias", "blk.weight"]):
            np.testing.assert_allclose(res[rank][n], total * (1 + i) / world)
    for n in res[0]:
        np.testing.assert_array_equal(res[0][n], res[1][n])
