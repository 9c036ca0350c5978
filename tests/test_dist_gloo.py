"""Multi-rank host logic on CPU (gloo, world_size 2): env sharding and the episode-statistics
all-reduce -- the only collective of the design."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from campx_b200 import dist as cxdist


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    r, w, _ = cxdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    lo, hi = cxdist.shard_range(10, rank, world)
    # per-rank stats block: episodes, sum, sumsq, len, max, -min, env_steps, reserved
    stats = torch.tensor([2.0 + rank, 10.0 * (rank + 1), 100.0, 50.0, 5.0 + rank, -1.0 + 3 * rank, 1000.0, 0.0],
                         dtype=torch.float64)
    cxdist.all_reduce_stats(stats)
    out.put((rank, lo, hi, stats.tolist()))
    dist.barrier()
    dist.destroy_process_group()


def test_shard_range_covers_everything_once():
    for n, w in ((10, 2), (7, 3), (1 << 20, 8), (5, 8)):
        spans = [cxdist.shard_range(n, r, w) for r in range(w)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
        assert max(h - l for l, h in spans) - min(h - l for l, h in spans) <= 1


def test_all_reduce_stats_world2_gloo():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(out.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [(r[1], r[2]) for r in res] == [(0, 5), (5, 10)]
    want = [5.0, 30.0, 200.0, 100.0, 6.0, 2.0, 2000.0, 0.0]
    for r in res:
        assert r[3] == want
    s = cxdist.summarize_stats(torch.tensor(want, dtype=torch.float64))
    assert s["episodes"] == 5 and s["return_mean"] == 6.0 and s["return_max"] == 6.0 and s["return_min"] == -2.0


def test_all_reduce_stats_is_identity_without_process_group():
    stats = torch.arange(8, dtype=torch.float64)
    assert cxdist.all_reduce_stats(stats.clone()).tolist() == stats.tolist()
    with pytest.raises(ValueError):
        cxdist.all_reduce_stats(torch.zeros(4, dtype=torch.float64))
