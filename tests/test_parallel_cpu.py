"""world_size-2 gloo test of the one collective on the path: disjoint-row sum all-reduce of
the per-image records, then identical ordered averaging on every rank."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import selfmask_b200 as S


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n_total, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(123)                       # every rank can regenerate the global truth
    counts = torch.from_numpy(rng.integers(0, 50000, (n_total, 2, 528), dtype=np.int32))
    sums = torch.from_numpy(rng.random((n_total, 2, 32)))
    a, b = S.shard_range(n_total, rank, world)
    full_c, full_s = S.allreduce_records(counts[a:b].clone(), sums[a:b].clone(), a, n_total)
    ok = torch.equal(full_c, counts) and torch.equal(full_s, sums)
    np.save(os.path.join(out_dir, f"ok_{rank}.npy"), np.array([ok, b - a]))
    dist.destroy_process_group()


def test_allreduce_records_is_an_exact_gather(tmp_path):
    world, n_total = 2, 11
    mp.spawn(_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    got = [np.load(tmp_path / f"ok_{r}.npy") for r in range(world)]
    assert all(g[0] == 1 for g in got)
    assert sum(int(g[1]) for g in got) == n_total
