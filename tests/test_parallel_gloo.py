"""N > 1 host logic on CPU: world_size-2 gloo process group (127.0.0.1 rendezvous) exercising clip sharding, the
ordered all-gather of results and the max-over-ranks timing reduction of audioeditingcode_b200.parallel."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, ws, port, n_clips, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from audioeditingcode_b200 import parallel as P
    assert P.shard_indices(n_clips, rank, ws) == list(range(rank, n_clips, ws))
    clips = [torch.full((1, 8, 4, 16), float(i)) for i in range(n_clips)]
    calls = []

    def edit(x):
        calls.append(int(x.flatten()[0]))
        return x * 2 + 1
    out = P.edit_clips(edit, clips)
    ok = all(torch.equal(out[i], clips[i] * 2 + 1) for i in range(n_clips))
    ok = ok and calls == list(range(rank, n_clips, ws))
    local_only = P.edit_clips(edit, clips, gather=False)
    ok = ok and all((local_only[i] is not None) == (i % ws == rank) for i in range(n_clips))
    mx = P.max_over_ranks(10.0 + rank)
    ok = ok and mx == 10.0 + (ws - 1)
    # row-ownership merge used by the timestep-sharded forward process
    full = torch.arange(7 * 3, dtype=torch.float32).reshape(7, 3) + 1
    mine = full.clone()
    owned = [i for i in range(7) if (i // 2) % ws == rank]
    mine[[i for i in range(7) if i not in owned]] = -99.0          # stale rows on this rank
    ok = ok and torch.equal(P.merge_owned_rows_(mine, owned), full)
    q.put((rank, ok))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 4, 1])
def test_clip_sharding_gloo_world2(n_clips):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, ws, port, n_clips, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


# ------------------------------------------------------------------ pc_drift: eigen-direction sharding + iterate all-gather
def _ag_worker(rank, ws, port, n_rows, q):
    """Host-side protocol of get_eigenvectors(group=...): round-robin row ownership, every rank contributes ONLY its owned
    rows of the [n_ev, D] iterate, one all-gather, rows re-assembled in direction order on every rank.  (The device math
    of the iteration has no CPU path; its 2-GPU NCCL check is tests/dist_gpu_check.py.)"""
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    from audioeditingcode_b200 import parallel as P
    full = torch.arange(n_rows * 5, dtype=torch.float32).reshape(n_rows, 5) * 1.5 + 1
    rows = P.shard_indices(n_rows, rank, ws)
    local = full[rows].clone() if rows else None
    got = P.allgather_rows(local, n_rows, full[0], group=dist.group.WORLD)
    ok = torch.equal(got, full)
    start = torch.full((3,), float(rank + 7))
    P.broadcast_(start, 0, dist.group.WORLD)                 # one random start for all ranks
    ok = ok and torch.equal(start, torch.full((3,), 7.0))
    q.put((rank, bool(ok)))
    dist.destroy_process_group()


@pytest.mark.parametrize("n_rows", [8, 3, 1])
def test_pc_drift_row_allgather_gloo_world2(n_rows):
    ws = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_ag_worker, args=(r, ws, port, n_rows, q)) for r in range(ws)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(ws)]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]
