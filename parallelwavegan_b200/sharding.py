"""Multi-GPU plumbing for the hot path: one process per GPU, no data-path
collective for inference -- utterances are partitioned round-robin (``i mod world``) -- and the
only exchanges are control-plane reductions of timings / counters (``torch.distributed``; NCCL on
GPUs, gloo in the CPU tests)."""
import torch


def partition(n_items, rank, world):
    """Indices of the utterances rank ``rank`` decodes: i mod world == rank (decode is embarrassingly
    parallel; the reference itself only ever decodes on one GPU, run.sh:147)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world {world}")
    return list(range(rank, n_items, world))


def reduce_stats(local_seconds, local_units, device=None, dist=None):
    """(max over ranks of seconds, sum over ranks of units).  Whole-job throughput =
    units / seconds with seconds the slowest rank's device time."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return float(local_seconds), float(local_units)
    t = torch.tensor([float(local_seconds)], dtype=torch.float64, device=device)
    u = torch.tensor([float(local_units)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(u, op=dist.ReduceOp.SUM)
    return float(t[0]), float(u[0])
