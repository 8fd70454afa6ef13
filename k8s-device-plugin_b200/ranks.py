"""Rank plumbing for the one-process-per-GPU launch (torchrun): barrier, max/sum over ranks.

The hot path shards per device with no exchange step, so there is NO data-path
collective; torch.distributed is used only to bracket timed regions and to reduce the timings
(NCCL on GPUs, gloo in the CPU tests).
"""
import os


class RankGroup:
    def __init__(self, backend=None, device=None):
        self.rank = int(os.environ.get("RANK", "0"))
        self.local_rank = int(os.environ.get("LOCAL_RANK", "0"))
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.backend = backend
        self.device = device
        self._dist = None
        if self.world > 1:
            import torch
            import torch.distributed as dist
            kw = {}
            if backend == "nccl":
                torch.cuda.set_device(self.local_rank)
                kw["device_id"] = torch.device("cuda", self.local_rank)
            dist.init_process_group(backend, **kw)
            self._dist = dist

    def _tensor(self, x):
        import torch
        dev = "cuda" if self.backend == "nccl" else "cpu"
        return torch.tensor([float(x)], dtype=torch.float64, device=dev)

    def barrier(self):
        if self.backend == "nccl":
            import torch
            torch.cuda.synchronize()
        if self._dist is not None:
            self._dist.barrier()
        if self.backend == "nccl":
            import torch
            torch.cuda.synchronize()

    def max(self, x):
        if self._dist is None:
            return float(x)
        t = self._tensor(x)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.MAX)
        return float(t.item())

    def sum(self, x):
        if self._dist is None:
            return float(x)
        t = self._tensor(x)
        self._dist.all_reduce(t, op=self._dist.ReduceOp.SUM)
        return float(t.item())

    def close(self):
        if self._dist is not None:
            self._dist.destroy_process_group()
            self._dist = None


def aggregate_throughput(rg: RankGroup, units_this_rank: float, seconds_this_rank: float):
    """Whole-job throughput = units all ranks processed / max-over-ranks time (weak scaling)."""
    total = rg.sum(units_this_rank)
    t = rg.max(seconds_this_rank)
    return total / t, t
