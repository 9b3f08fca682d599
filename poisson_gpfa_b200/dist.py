"""Trial sharding across GPUs.

Trials are conditionally independent given (C, d, tau) (funs/inference.py:94), so rank g of G owns a
contiguous block of trials for the whole fit and the E-step needs no communication.  The exchange
steps are sums of small sufficient statistics: the objective sum, PautoSum (q x T x T) once per
E-step, and the per-neuron (cost, gradient, Hessian) block once per M-step Newton iteration.
They go through one NCCL all-reduce each (torch.distributed is the plumbing); every rank then solves
the identical tiny M-step redundantly, so no broadcast is needed.  With a single process the
reducer is the identity.
"""
import os

import torch


def shard_bounds(num_trials, world_size, rank):
    """Contiguous block [lo, hi) of trials owned by `rank`; sizes differ by at most one."""
    base, rem = divmod(int(num_trials), int(world_size))
    lo = rank * base + min(rank, rem)
    hi = lo + base + (1 if rank < rem else 0)
    return lo, hi


def batch_shard(batch_indices, num_trials, world_size, rank):
    """Positions (within the mini-batch) and local trial offsets of the mini-batch members owned by `rank`."""
    lo, hi = shard_bounds(num_trials, world_size, rank)
    pos = [i for i, t in enumerate(batch_indices) if lo <= int(t) < hi]
    return pos, [int(batch_indices[i]) - lo for i in pos]


class Reducer:
    """Sum-reductions over the data-parallel group; identity when torch.distributed is not initialised."""

    def __init__(self, group=None):
        import torch.distributed as td
        self._td = td
        self.active = td.is_available() and td.is_initialized()
        self.group = group
        self.world_size = td.get_world_size(group) if self.active else 1
        self.rank = td.get_rank(group) if self.active else 0
        self.n_allreduce = 0

    def sum_tensor(self, t):
        if self.active and self.world_size > 1:
            self._td.all_reduce(t, op=self._td.ReduceOp.SUM, group=self.group)
            self.n_allreduce += 1
        return t

    def sum_scalar(self, v):
        if not (self.active and self.world_size > 1):
            return v
        dev = "cuda" if torch.cuda.is_available() and self._td.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        self._td.all_reduce(t, op=self._td.ReduceOp.SUM, group=self.group)
        self.n_allreduce += 1
        return float(t.item())

    def max_scalar(self, v):
        if not (self.active and self.world_size > 1):
            return v
        dev = "cuda" if torch.cuda.is_available() and self._td.get_backend(self.group) == "nccl" else "cpu"
        t = torch.tensor([float(v)], dtype=torch.float64, device=dev)
        self._td.all_reduce(t, op=self._td.ReduceOp.MAX, group=self.group)
        return float(t.item())


def init_from_env():
    """Initialise torch.distributed from torchrun's environment (one process per GPU, NCCL)."""
    import torch.distributed as td
    if td.is_initialized():
        return Reducer()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world <= 1:
        return Reducer()
    local = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
    if torch.cuda.is_available():
        torch.cuda.set_device(local)
        td.init_process_group(backend="nccl", device_id=torch.device("cuda", local))
    else:
        td.init_process_group(backend="gloo")
    return Reducer()
