"""Host-side plumbing of the layer-split pipeline (one process per GPU): which layers a rank owns, how the NCCL unique id
reaches every rank, how the per-rank timings become one number. torch.distributed is plumbing only — the data path is
b200_pipeline_* (one ncclSend/ncclRecv of the residual stream per stage boundary, booster_b200/csrc/engine.cu).

Mirrors the reference's layer split (cpp/src/llama.cpp:5932-5968): normalised cumulative proportions and an upper_bound
on il / (n_layer + 1); with equal proportions this is the contiguous equal split bench.py uses."""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple


def split_layers(n_layer: int, proportions: Sequence[float]) -> List[int]:
    """layer -> stage, the rule of booster_b200/csrc/bridge.cpp::split_layers (and llm_load_tensors)"""
    if not proportions or sum(proportions) <= 0:
        raise ValueError("proportions must contain a positive entry")
    import numpy as np
    cum = np.cumsum(np.asarray(proportions, dtype=np.float32), dtype=np.float32)
    cum = cum / cum[-1]
    act = n_layer + 1
    out = []
    for il in range(n_layer):
        f = np.float32(il) / np.float32(act)
        d = 0
        while d + 1 < len(cum) and not (f < cum[d]):
            d += 1
        out.append(d)
    return out


def stage_range(n_layer: int, rank: int, world: int, proportions: Optional[Sequence[float]] = None) -> Tuple[int, int]:
    """[begin, end) of the layers rank `rank` owns; ranks without layers get an empty range (begin == end)"""
    if not 0 <= rank < world:
        raise ValueError("rank out of range")
    if proportions is None:
        return rank * n_layer // world, (rank + 1) * n_layer // world
    if len(proportions) != world:
        raise ValueError("one proportion per rank")
    dev = split_layers(n_layer, proportions)
    mine = [i for i, d in enumerate(dev) if d == rank]
    if not mine:
        return 0, 0
    assert mine == list(range(mine[0], mine[-1] + 1)), "stages are contiguous by construction"
    return mine[0], mine[-1] + 1


def share_unique_id(dist, make_id) -> bytes:
    """rank 0 creates the 128-byte ncclUniqueId (b200_comm_unique_id), every rank receives it over the CPU group"""
    uid = [make_id() if dist.get_rank() == 0 else None]
    dist.broadcast_object_list(uid, src=0)
    if not isinstance(uid[0], (bytes, bytearray)) or len(uid[0]) != 128:
        raise RuntimeError("unique id must be 128 bytes")
    return bytes(uid[0])


def max_over_ranks(dist, seconds: float, device=None) -> float:
    """the timing rule of bench.py: every rank brackets the same steps with barriers, the job's time is the slowest rank's"""
    import torch
    t = torch.tensor([seconds], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(dist, n: int, device=None) -> int:
    import torch
    t = torch.tensor([n], dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(t.item())


def exchange_inbox_handles(dist, my_handle: bytes) -> Tuple[bytes, bytes]:
    """every rank contributes its 64-byte inbox handle (b200_p2p_handle); returns (the next rank's, rank 0's) — what
    b200_p2p_connect needs. The last rank's `next` is its own handle (unused)."""
    if len(my_handle) != 64:
        raise RuntimeError("an inbox handle is 64 bytes")
    world, rank = dist.get_world_size(), dist.get_rank()
    handles = [None] * world
    dist.all_gather_object(handles, my_handle)
    if any(not isinstance(h, (bytes, bytearray)) or len(h) != 64 for h in handles):
        raise RuntimeError("bad inbox handle from a peer")
    return bytes(handles[min(rank + 1, world - 1)]), bytes(handles[0])


def connect_peer_handoff(dist, ctx, rank: int, world: int) -> bool:
    """set up the NVLink peer hand-off on every rank, or on none: a rank that cannot export / map an inbox (no peer access
    between two devices, IPC refused) makes the whole group fall back to ncclSend / ncclRecv. Returns what was decided."""
    ok, handle = 1, bytes(64)
    try:
        handle = ctx.p2p_handle()
    except Exception:
        ok = 0
    try:
        nxt, first = exchange_inbox_handles(dist, handle)
        if ok:
            ctx.p2p_connect(rank, world, nxt, first)
    except Exception:
        ok = 0
    all_ok = min_over_ranks(dist, ok)
    if not all_ok:
        try:
            ctx.p2p_disable()
        except Exception:
            pass
    return bool(all_ok)


def min_over_ranks(dist, v: int) -> int:
    import torch
    t = torch.tensor([int(v)], dtype=torch.int64)
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    return int(t.item())
