"""Batch sharding across GPUs (no reference analogue -- the reference is single-process).

Every batch item is an independent NLS problem that shares one symbolic structure, so the batch dimension shards with
no data-path collective.  The only batch-global decisions of the reference loop are reproduced with one tiny
all-reduce per LM iteration:
  * all-rejected retry  (nonlinear_least_squares.py:181-188, 358)  -> SUM of [#rejected, #items]
  * mean-error / all-converged tests (nonlinear_optimizer.py:111, nonlinear_least_squares.py:202) -> SUM of [sum|err|, #converged, #items]
Works with any torch.distributed backend (NCCL on the GPUs, gloo in the CPU tests).
"""
from typing import Optional, Tuple

import torch


def batch_shard(total: int, rank: int, world: int) -> slice:
    """Contiguous, balanced slice of the batch owned by `rank` (first `total % world` ranks get one extra item)."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside world of size {world}")
    base, extra = divmod(total, world)
    start = rank * base + min(rank, extra)
    return slice(start, start + base + (1 if rank < extra else 0))


def reduce_counts(stats: torch.Tensor, group=None) -> torch.Tensor:
    """In-place SUM all-reduce of a small stats tensor; identity when not distributed."""
    if group is not None:
        import torch.distributed as dist
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def all_rejected(num_rejected_local: int, num_items_local: int, group=None, device: Optional[torch.device] = None) -> bool:
    """True iff every batch item on every rank was rejected."""
    t = torch.tensor([num_rejected_local, num_items_local], dtype=torch.int64, device=device)
    reduce_counts(t, group)
    return int(t[0]) == int(t[1])


def global_mean_abs_error(err_local: torch.Tensor, group=None) -> Tuple[float, int]:
    """(mean over the GLOBAL batch of |err|, global batch size) -- the quantity nonlinear_optimizer.py:111 thresholds."""
    t = torch.stack([err_local.abs().sum().double(), torch.tensor(float(err_local.numel()), dtype=torch.float64, device=err_local.device)])
    reduce_counts(t, group)
    return float(t[0] / t[1]), int(t[1])


def any_rank_true(flag_local: bool, group=None, device: Optional[torch.device] = None) -> bool:
    """True iff `flag_local` is True on ANY rank (every rank must call it): exit decisions of the sharded loop -- a failed linear
    solve on one rank ends the loop on all of them, so the per-iteration collectives stay matched."""
    if group is None:
        return bool(flag_local)
    t = torch.tensor([1 if flag_local else 0], dtype=torch.int64, device=device)
    reduce_counts(t, group)
    return int(t[0]) > 0
