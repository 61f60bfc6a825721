"""World-size-2 gloo test (CPU) of the multi-GPU host logic: stream sharding, max-over-ranks timing and
whole-job throughput aggregation (bench.py's N>1 path)."""
import os
import socket

import pytest
import torch.multiprocessing as mp

from nubovca import shard


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    mine = shard.streams_of_rank(9, world, rank)
    ms = 10.0 + 5.0 * rank                       # rank 1 is the slow one
    value, ms_max = shard.aggregate_throughput(len(mine) * 30.0, ms, dist)
    mx = shard.reduce_max([ms, rank], dist)
    dist.barrier()
    q.put((rank, mine, value, ms_max, mx))
    dist.destroy_process_group()


def test_two_ranks_gloo():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, s0, v0, m0, x0), (_, s1, v1, m1, x1) = res
    assert sorted(s0 + s1) == list(range(9)) and not set(s0) & set(s1)       # every stream on exactly one GPU
    assert s0 == [0, 2, 4, 6, 8] and s1 == [1, 3, 5, 7]
    assert m0 == m1 == 15.0 and x0 == x1 == [15.0, 1.0]                        # the slowest rank's time
    assert v0 == v1 == pytest.approx(9 * 30.0 / 0.015)                         # whole-job units / slowest time


def test_single_process_is_identity():
    assert shard.streams_of_rank(5, 1, 0) == [0, 1, 2, 3, 4]
    assert shard.aggregate_throughput(100.0, 50.0) == (2000.0, 50.0)
    with pytest.raises(ValueError):
        shard.streams_of_rank(5, 2, 2)
