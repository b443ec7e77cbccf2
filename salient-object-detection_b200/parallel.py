"""Batch-sharded data parallelism: one process per GPU, images partitioned by contiguous
index ranges, weights replicated, and ONE collective per exchange: the per-image integer count rows (histograms, counts at
the thresholds, centroids) and the float64 moment rows, packed into a single int32 record per evaluated mask and gathered
over NVLink (`all_gather_into_tensor`; rows are disjoint per rank, so the gather is exactly the disjoint-row sum all-reduce
of SURVEY §8e without moving 7/8 zeros).  Ratios and the ordered running means are then formed identically on every rank.
The reference has no distributed code at all (§2.2); this is new.

`RecordExchange` keeps the collective off the critical path: preallocated send / receive buffers (two slots), the collective
is issued asynchronously (NCCL runs it on its own stream, after the evaluation kernels that produced the rows) and only
waited for when the gathered records are read — the next step's encoder runs meanwhile.
"""
from typing import List, Optional, Tuple

import torch
import torch.distributed as dist

from ._lib import MCOUNT_STRIDE, MSUM_STRIDE

ROW_WORDS = MCOUNT_STRIDE + 2 * MSUM_STRIDE        # int32 words of one packed mask record: 528 counts + 32 doubles


def shard_range(n_items: int, rank: int, world_size: int) -> Tuple[int, int]:
    """Contiguous [start, stop) of rank `rank`: ceil(n/W) per rank, the tail ranks may get fewer / none."""
    per = -(-n_items // world_size)
    start = min(rank * per, n_items)
    return start, min(start + per, n_items)


def _world(group=None) -> int:
    return dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1


class RecordExchange:
    """Gathers every rank's per-image records ([n_local, 2, 528] int32 + [n_local, 2, 32] float64) on every rank.

    `capacity` = the largest number of local images any rank posts at once (ragged shards are padded to it; the pad rows are
    dropped on the receiving side from the row counts each rank sends along in the record header)."""

    SLOTS = 2

    def __init__(self, capacity: int, device, group: Optional[dist.ProcessGroup] = None):
        self.capacity, self.group, self.device = int(capacity), group, torch.device(device)
        self.world = _world(group)
        # slot layout: [capacity + 1, 2, ROW_WORDS] int32 — row 0 is a header (word 0 = number of valid rows)
        shape = (self.capacity + 1, 2, ROW_WORDS)
        self.send = [torch.zeros(shape, dtype=torch.int32, device=self.device) for _ in range(self.SLOTS)]
        # receive side: the ranks' slots concatenated along dim 0 (the layout both NCCL and gloo accept)
        self.recv = [torch.zeros((self.world * shape[0],) + shape[1:], dtype=torch.int32, device=self.device) for _ in range(self.SLOTS)]
        self.work: List[Optional[object]] = [None] * self.SLOTS
        self.posted = 0

    def post(self, m_counts: torch.Tensor, m_sums: torch.Tensor) -> int:
        """Pack this rank's rows into the next slot and start the gather (asynchronous).  Returns the slot ticket."""
        n = int(m_counts.shape[0])
        if n > self.capacity:
            raise ValueError(f"{n} rows exceed the exchange capacity {self.capacity}")
        slot = self.posted % self.SLOTS
        if self.work[slot] is not None:          # the slot's previous collective must have finished before its buffers are reused
            self.work[slot].wait()
            self.work[slot] = None
        send = self.send[slot]
        send[0, 0, 0] = n
        if n:
            send[1:n + 1, :, :MCOUNT_STRIDE].copy_(m_counts.reshape(n, 2, MCOUNT_STRIDE))
            send[1:n + 1, :, MCOUNT_STRIDE:].copy_(m_sums.reshape(n, 2, MSUM_STRIDE).contiguous().view(torch.int32))
        if self.world > 1:
            self.work[slot] = dist.all_gather_into_tensor(self.recv[slot], send, group=self.group, async_op=True)
        else:
            self.recv[slot].copy_(send)
        self.posted += 1
        return slot

    def collect(self, slot: int) -> Tuple[torch.Tensor, torch.Tensor]:
        """Wait for the slot's gather and return (counts [n_total, 2, 528] int32, sums [n_total, 2, 32] float64) in rank order."""
        if self.work[slot] is not None:
            self.work[slot].wait()
            self.work[slot] = None
        recv = self.recv[slot].view(self.world, self.capacity + 1, 2, ROW_WORDS)
        ns = recv[:, 0, 0, 0].tolist()           # rows per rank (one small device → host read at collection time)
        if all(n == self.capacity for n in ns):  # uniform shards: a single strided view, no gather kernel
            body = recv[:, 1:].reshape(self.world * self.capacity, 2, ROW_WORDS)
        else:
            body = torch.cat([recv[r, 1:n + 1] for r, n in enumerate(ns)]) if sum(ns) else recv[0, 1:1]
        counts = body[..., :MCOUNT_STRIDE].contiguous()
        sums = body[..., MCOUNT_STRIDE:].contiguous().view(torch.float64)
        return counts, sums


def gather_records(local_counts: torch.Tensor, local_sums: torch.Tensor,
                   group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """One-shot form: every rank's rows (contiguous, rank-ordered shards as `shard_range` deals them; a rank may hold none) →
    the records of the whole sweep, in dataset order, on every rank (NCCL over NVLink on GPUs, gloo in the CPU tests)."""
    world = _world(group)
    n_local = int(local_counts.shape[0])
    cap = n_local
    if world > 1:
        t = torch.tensor([n_local], dtype=torch.int64, device=local_counts.device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
        cap = int(t.item())
    ex = RecordExchange(max(cap, 1), local_counts.device, group)
    return ex.collect(ex.post(local_counts.reshape(n_local, 2, MCOUNT_STRIDE), local_sums.reshape(n_local, 2, MSUM_STRIDE)))


def allreduce_records(local_counts: torch.Tensor, local_sums: torch.Tensor, start: int, n_total: int,
                      group: Optional[dist.ProcessGroup] = None) -> Tuple[torch.Tensor, torch.Tensor]:
    """The exchange for callers that know the sweep size: local rows are images [start, start + n_local) of n_total, dealt by
    `shard_range` — the padded per-rank capacity is then ceil(n_total / world) and the gather is the ONLY collective issued."""
    world, n_local = _world(group), int(local_counts.shape[0])
    cap = max(-(-n_total // world), 1)
    if n_local > cap:
        raise ValueError(f"{n_local} local rows exceed ceil(n_total / world) = {cap}: shards must come from shard_range")
    ex = RecordExchange(cap, local_counts.device, group)
    counts, sums = ex.collect(ex.post(local_counts.reshape(n_local, 2, MCOUNT_STRIDE), local_sums.reshape(n_local, 2, MSUM_STRIDE)))
    if counts.shape[0] != n_total:
        raise ValueError(f"gathered {counts.shape[0]} rows, expected {n_total}")
    return counts, sums
