"""Batch-sharded data parallelism: one process per GPU, images partitioned by contiguous
index ranges, weights replicated, and ONE collective — a sum all-reduce of the per-image integer count rows
(disjoint rows per rank ⇒ an exact gather) plus the double-precision moment rows (adding zeros is exact).
Ratios and the ordered running means are then formed identically on every rank.
The reference has no distributed code at all (§2.2); this is new.
"""
from typing import Optional, Tuple

import torch
import torch.distributed as dist


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank `rank`: ceil(n/W) per rank, the tail ranks may get fewer / none."""
    per = -(-n_items // world_size)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def allreduce_records(local_counts: torch.Tensor, local_sums: torch.Tensor, start: int, n_total: int,
                      group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """local_counts int32 [n_local, ...], local_sums float64 [n_local, ...] for images [start, start+n_local)
    → the full [n_total, ...] records on every rank (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    full_c = torch.zeros((n_total,) + tuple(local_counts.shape[1:]), dtype=local_counts.dtype, device=local_counts.device)
    full_s = torch.zeros((n_total,) + tuple(local_sums.shape[1:]), dtype=local_sums.dtype, device=local_sums.device)
    n_local = local_counts.shape[0]
    full_c[start:start + n_local] = local_counts
    full_s[start:start + n_local] = local_sums
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(full_c, op=dist.ReduceOp.SUM, group=group)
        dist.all_reduce(full_s, op=dist.ReduceOp.SUM, group=group)
    return full_c, full_s
