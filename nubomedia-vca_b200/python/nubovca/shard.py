"""Stream -> GPU sharding: video streams are independent, so a box shards them by stream
with no data-path collective; the only cross-rank traffic is the benchmark's barrier and max-reduce."""
from __future__ import annotations


def streams_of_rank(n_streams: int, world: int, rank: int) -> list[int]:
    """gpu = stream_id mod n_gpus, sticky for the stream's lifetime (temporal state stays on its GPU)."""
    if world < 1 or not 0 <= rank < world:
        raise ValueError("bad world/rank")
    return [s for s in range(n_streams) if s % world == rank]


def reduce_max(values, dist=None, device=None):
    """Max over ranks of a list of floats (timings are reported as the slowest rank's)."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.tolist()


def reduce_sum(values, dist=None, device=None):
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(v) for v in values]
    import torch
    t = torch.tensor([float(v) for v in values], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.tolist()


def aggregate_throughput(units_per_rank: float, ms_per_rank: float, dist=None, device=None) -> tuple[float, float]:
    """Whole-job units/s = (units all ranks processed) / (slowest rank's time).  Returns (units/s, ms)."""
    total = reduce_sum([units_per_rank], dist, device)[0]
    ms = reduce_max([ms_per_rank], dist, device)[0]
    return total / (ms * 1e-3), ms
