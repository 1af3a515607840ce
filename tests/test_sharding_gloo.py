"""N>1 host logic on CPU: world_size-2 gloo process group, pair partitioning, the pose-record gather (the
path's only collective) and the max-over-ranks timing reduction."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from gfpl import _structs as S
from gfpl import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _records(lo, hi):
    r = np.zeros(hi - lo, S.pose_result_dt)
    r["err_norm"] = np.arange(lo, hi)
    r["n_inliers"] = np.arange(lo, hi) * 3
    r["DT"][:, 0] = np.arange(lo, hi) + 0.5
    return r


def _worker(rank, world, port, n_pairs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard.partition(n_pairs, world, rank)
        rec = _records(lo, hi)
        allrec = shard.gather_pose_records_ragged(rec, world)
        ok = len(allrec) == n_pairs and np.array_equal(allrec["err_norm"], np.arange(n_pairs)) \
            and np.array_equal(allrec["n_inliers"], np.arange(n_pairs) * 3)
        # equal-sized fast path (what bench.py uses with 8192 pairs per rank)
        eq = torch.from_numpy(_records(rank * 5, rank * 5 + 5).view(np.uint8).reshape(-1).copy())
        g = shard.gather_pose_records(eq, world).numpy().view(S.pose_result_dt)
        ok = ok and np.array_equal(g["err_norm"], np.arange(5 * world))
        t = shard.max_over_ranks(10.0 + rank, world)
        ok = ok and t == 10.0 + world - 1
        q.put((rank, bool(ok)))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_pairs", [8192, 37])
def test_world2_gather(n_pairs):
    ctxm = mp.get_context("spawn")
    q = ctxm.Queue()
    port = _free_port()
    ps = [ctxm.Process(target=_worker, args=(r, 2, port, n_pairs, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = [q.get(timeout=120) for _ in ps]
    for p in ps:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_partition_covers_everything():
    for n, w in ((8192, 8), (37, 4), (3, 8), (0, 2)):
        spans = [shard.partition(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        sizes = [hi - lo for lo, hi in spans]
        assert max(sizes) - min(sizes) <= 1
