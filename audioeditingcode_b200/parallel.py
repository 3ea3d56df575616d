"""Multi-GPU execution of the hot path: the path shards by CLIP — one process per GPU
(`torchrun`, backend nccl over NVLink / NVSwitch), weights replicated, every rank edits its own clips with the
single-GPU code.  There is no data-path collective; the only communication is the final gather of the edited
latents (and the max-over-ranks reduction of timings in bench.py).  The reverse process of one clip is sequential in
t and does not shard (replicas only); the reference itself has no multi-GPU code on this path.
"""
from __future__ import annotations

from typing import Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def world(group=None) -> tuple:
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(group), dist.get_world_size(group)
    return 0, 1


def broadcast_(t: torch.Tensor, src: int = 0, group=None) -> torch.Tensor:
    """In-place broadcast from group rank `src` (no-op without a process group)."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.broadcast(t, src=dist.get_global_rank(group, src) if group is not None else src, group=group)
    return t


def allgather_rows(local: Optional[torch.Tensor], n_rows: int, like_row: torch.Tensor, group=None) -> torch.Tensor:
    """Assemble a `[n_rows, ...]` tensor whose rows were computed on different ranks (round-robin ownership, see
    shard_indices): every rank contributes ONLY its owned rows — one all-gather of ceil(n_rows / world) rows per rank,
    1/world of the bytes of a zero-padded sum-all-reduce — and every rank ends with bit-identical rows (the pc_drift
    iterate exchange named by BASELINE configs[3])."""
    rank, ws = world(group)
    if ws == 1:
        return local
    per = (n_rows + ws - 1) // ws
    send = torch.zeros((per, *like_row.shape), dtype=like_row.dtype, device=like_row.device)
    mine = shard_indices(n_rows, rank, ws)
    if mine:
        send[:len(mine)] = local.to(send.dtype)
    recv = torch.empty((ws * per, *like_row.shape), dtype=like_row.dtype, device=like_row.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    # row i lives at rank i % ws, slot i // ws
    idx = torch.as_tensor([(i % ws) * per + i // ws for i in range(n_rows)], device=like_row.device)
    return recv.index_select(0, idx)


def shard_indices(n_items: int, rank: int, world_size: int) -> List[int]:
    """Round-robin assignment: item i -> rank i % world_size (keeps per-rank counts within 1 of each other)."""
    return list(range(rank, n_items, world_size))


def edit_clips(edit_fn: Callable[[torch.Tensor], torch.Tensor], clips: Sequence[torch.Tensor],
               gather: bool = True, group=None) -> List[Optional[torch.Tensor]]:
    """Run `edit_fn` (e.g. inversion_forward_process + inversion_reverse_process on one clip latent) over `clips`,
    sharded across the ranks of the default process group.  Returns the results in clip order — on every rank if
    `gather`, else only this rank's entries (others None).  All clips must produce equally shaped results."""
    rank, ws = world()
    mine = shard_indices(len(clips), rank, ws)
    local = {i: edit_fn(clips[i]) for i in mine}
    out: List[Optional[torch.Tensor]] = [None] * len(clips)
    for i, t in local.items():
        out[i] = t
    if ws == 1 or not gather or not clips:
        return out
    per_rank = (len(clips) + ws - 1) // ws
    ref = next(iter(local.values())) if local else None
    shape = [None, None]
    if rank == 0:
        shape = [tuple(ref.shape), str(ref.dtype)]
    dist.broadcast_object_list(shape, src=0, group=group)
    shp = shape[0]
    dtype = ref.dtype if ref is not None else getattr(torch, shape[1].split(".")[-1])
    device = ref.device if ref is not None else (torch.device("cuda", torch.cuda.current_device())
                                                 if dist.get_backend(group) == "nccl" else torch.device("cpu"))
    buf = torch.zeros((per_rank, *shp), dtype=dtype, device=device)
    for k, i in enumerate(mine):
        buf[k] = local[i]
    bufs = [torch.empty_like(buf) for _ in range(ws)]
    dist.all_gather(bufs, buf, group=group)
    for r in range(ws):
        for k, i in enumerate(shard_indices(len(clips), r, ws)):
            out[i] = bufs[r][k]
    return out


def merge_owned_rows_(t: torch.Tensor, owned: Sequence[int], group=None) -> torch.Tensor:
    """In place: afterwards every rank holds every row of `t` (dim 0) from the rank that owns it, bit-exactly.  Rows must
    be owned by exactly one rank; ownership is arbitrary (the timestep chunks of the sharded forward process), so the
    ranks first exchange their index lists (n_rows int64 each), then ONE all-gather of max-owned-count rows per rank —
    1/world of the bytes of a zero-padded sum-all-reduce of the full tensor."""
    if not (dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1):
        return t
    ws = dist.get_world_size(group)
    n = t.shape[0]
    idx = torch.full((n,), -1, dtype=torch.int64, device=t.device)
    if len(owned):
        idx[:len(owned)] = torch.as_tensor(list(owned), dtype=torch.int64, device=t.device)
    all_idx = torch.empty((ws * n,), dtype=torch.int64, device=t.device)      # concatenated layout (gloo and nccl)
    dist.all_gather_into_tensor(all_idx, idx, group=group)
    all_idx = all_idx.view(ws, n)
    counts = (all_idx >= 0).sum(1)
    per = int(counts.max().item())
    if int(counts.sum().item()) != n or per == 0:
        raise ValueError(f"merge_owned_rows_: {int(counts.sum().item())} owned rows over the group for a tensor of {n} rows")
    send = torch.zeros((per, *t.shape[1:]), dtype=t.dtype, device=t.device)
    if len(owned):
        send[:len(owned)] = t[idx[:len(owned)]]
    recv = torch.empty((ws * per, *t.shape[1:]), dtype=t.dtype, device=t.device)
    dist.all_gather_into_tensor(recv, send, group=group)
    recv = recv.view(ws, per, *t.shape[1:])
    valid = all_idx[:, :per] >= 0
    t[all_idx[:, :per][valid]] = recv[valid]
    return t


def max_over_ranks(value: float, device=None) -> float:
    """Device-timed durations are reduced with MAX over ranks (never wall-clock, never mean)."""
    rank, ws = world()
    if ws == 1:
        return float(value)
    dev = device if device is not None else (torch.device("cuda", torch.cuda.current_device())
                                             if dist.get_backend() == "nccl" else torch.device("cpu"))
    t = torch.tensor([float(value)], device=dev, dtype=torch.float64)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
