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


def _sweep_worker(rank, world, port, n_total, out_dir):
    """BASELINE.json configs[4] in miniature: every rank evaluates its contiguous shard of a dataset sweep (here: records of
    small synthetic masks built with the numpy emulation of the GPU record layout), all-reduces the per-image rows once, and
    must arrive at exactly the 14 averages a single process computes over the whole sweep in dataset order."""
    from tests.helpers import numpy_record
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def record(i):                                        # image i of the sweep: 2 evaluated masks (selected, upper bound)
        rng = np.random.default_rng(1000 + i)
        gt = np.zeros((24, 20), bool)
        if i % 7 != 3:                                    # every 7th ground truth is empty (reference edge case)
            y0, x0 = rng.integers(0, 12), rng.integers(0, 10)
            gt[y0:y0 + rng.integers(4, 12), x0:x0 + rng.integers(4, 10)] = True
        recs = [numpy_record(rng.random((24, 20), dtype=np.float32), gt) for _ in range(2)]
        return np.stack([r[0] for r in recs]), np.stack([r[1] for r in recs])

    a, b = S.shard_range(n_total, rank, world)
    mine = [record(i) for i in range(a, b)]
    lc = torch.from_numpy(np.stack([m[0] for m in mine]).astype(np.int32)) if mine else torch.zeros((0, 2, 528), dtype=torch.int32)
    ls = torch.from_numpy(np.stack([m[1] for m in mine])) if mine else torch.zeros((0, 2, 32), dtype=torch.float64)
    full_c, full_s = S.allreduce_records(lc, ls, a, n_total)
    res = S.summarize(full_c.numpy(), full_s.numpy())
    if rank == 0:                                          # single-process truth over the whole sweep
        allr = [record(i) for i in range(n_total)]
        ref = S.summarize(np.stack([m[0] for m in allr]).astype(np.int32), np.stack([m[1] for m in allr]))
        same = all((np.isnan(res[k]) and np.isnan(ref[k])) or res[k] == ref[k] for k in ref)
        np.save(os.path.join(out_dir, "sweep_ok.npy"), np.array([same, len(ref)]))
    keys = sorted(res)
    mine_vals = torch.tensor([res[k] for k in keys], dtype=torch.float64)
    gathered = [torch.zeros_like(mine_vals) for _ in range(world)]
    dist.all_gather(gathered, mine_vals)
    agree = all(torch.equal(torch.nan_to_num(g, nan=-1.0), torch.nan_to_num(gathered[0], nan=-1.0)) for g in gathered)
    np.save(os.path.join(out_dir, f"agree_{rank}.npy"), np.array([agree]))
    dist.destroy_process_group()


def test_sharded_sweep_gives_the_single_process_averages(tmp_path):
    world, n_total = 2, 37                                 # odd count: the last rank gets the shorter shard
    mp.spawn(_sweep_worker, args=(world, _free_port(), n_total, str(tmp_path)), nprocs=world, join=True)
    ok = np.load(tmp_path / "sweep_ok.npy")
    assert ok[0] == 1 and ok[1] == 14
    assert all(np.load(tmp_path / f"agree_{r}.npy")[0] == 1 for r in range(world))


def test_shard_ranges_of_the_duts_te_sweep():
    """5019 images over 8 ranks (configs[4]): contiguous, disjoint, complete; ceil(n / W) = 628 per rank, the last rank takes the remaining 623."""
    spans = [S.shard_range(5019, r, 8) for r in range(8)]
    assert spans[0][0] == 0 and spans[-1][1] == 5019
    assert all(spans[i][1] == spans[i + 1][0] for i in range(7))
    sizes = [b - a for a, b in spans]
    assert max(sizes) == 628 and min(sizes) >= 623


def _exchange_worker(rank, world, port, n_total, out_dir):
    """RecordExchange: preallocated two-slot buffers, several exchanges in flight, ragged and EMPTY shards (1 image over 2 ranks)."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ok = True
    cap = -(-n_total // world)
    ex = S.RecordExchange(max(cap, 1), "cpu")
    tickets, truths = [], []
    for step in range(5):                                  # more posts than slots: slot reuse must wait for the earlier gather
        rng = np.random.default_rng(7 + step)
        counts = torch.from_numpy(rng.integers(-5, 50000, (n_total, 2, 528), dtype=np.int32))
        sums = torch.from_numpy(rng.standard_normal((n_total, 2, 32)))
        sums[0, 0, 3] = float("nan")                       # NaN moments (the reference's NaN S-measure cases) must travel bit-exactly
        a, b = S.shard_range(n_total, rank, world)
        tickets.append(ex.post(counts[a:b].clone(), sums[a:b].clone()))
        truths.append((counts, sums))
        if step >= 1:                                      # collect one step late, as an overlapped caller does
            c, s_ = ex.collect(tickets[step - 1])
            tc, ts = truths[step - 1]
            ok = ok and torch.equal(c, tc) and torch.equal(s_.view(torch.int64), ts.view(torch.int64))
    c, s_ = ex.collect(tickets[-1])
    ok = ok and torch.equal(c, truths[-1][0]) and torch.equal(s_.view(torch.int64), truths[-1][1].view(torch.int64))
    g_c, g_s = S.gather_records(truths[0][0][slice(*S.shard_range(n_total, rank, world))], truths[0][1][slice(*S.shard_range(n_total, rank, world))])
    ok = ok and torch.equal(g_c, truths[0][0])
    np.save(os.path.join(out_dir, f"ex_{rank}.npy"), np.array([ok]))
    dist.destroy_process_group()


def test_record_exchange_ragged_and_empty_shards(tmp_path):
    for n_total in (11, 1, 8):                             # ragged (6 + 5), one rank empty (1 + 0), uniform (4 + 4)
        d = tmp_path / str(n_total)
        d.mkdir()
        mp.spawn(_exchange_worker, args=(2, _free_port(), n_total, str(d)), nprocs=2, join=True)
        assert all(np.load(d / f"ex_{r}.npy")[0] == 1 for r in range(2)), n_total
