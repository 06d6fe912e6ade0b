"""Stream -> GPU sharding and cross-rank reporting for the multi-GPU path.

Independent wideband streams (one per SDR source, rc_frontend/receiver.py:67-70 /
systemd/radiocapture-channelizer@.service:11) shard embarrassingly: stream s runs on rank s mod G, no
data-path collective.  torch.distributed is only used for the barrier and for reducing the timing
(MAX over ranks) and counters (SUM) - it is imported lazily and only when WORLD_SIZE > 1, so the
single-GPU product path stays torch-free.
"""
import os


def world_from_env():
    return (int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("RANK", "0")),
            int(os.environ.get("LOCAL_RANK", "0")))


def assign_streams(n_streams, world_size, rank):
    """Round-robin: stream s -> rank s mod world_size.  Returns the stream ids this rank owns."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("bad world_size / rank")
    return [s for s in range(n_streams) if s % world_size == rank]


def streams_per_rank(n_streams, world_size):
    return [len(assign_streams(n_streams, world_size, r)) for r in range(world_size)]


class Reducer(object):
    """Barrier + MAX / SUM / MIN reductions over ranks (gloo on CPU tensors, nccl on cuda tensors)."""

    def __init__(self, dist=None, device=None):
        self.dist = dist
        self.device = device

    def _reduce(self, v, op):
        if self.dist is None:
            return float(v)
        import torch
        t = torch.tensor([float(v)], dtype=torch.float64, device=self.device or "cpu")
        self.dist.all_reduce(t, op=getattr(self.dist.ReduceOp, op))
        return float(t.item())

    def max(self, v):
        return self._reduce(v, "MAX")

    def min(self, v):
        return self._reduce(v, "MIN")

    def sum(self, v):
        return self._reduce(v, "SUM")

    def barrier(self):
        if self.dist is not None:
            self.dist.barrier()


def whole_job_rate(samples_this_rank, elapsed_ms_this_rank, reducer):
    """value = units all ranks processed / MAX-over-ranks time  (bench.py contract)."""
    total = reducer.sum(samples_this_rank)
    ms = reducer.max(elapsed_ms_this_rank)
    return total / (ms * 1e-3), ms
