"""Data-parallel serving across the GPUs of one box: utterances are independent, so the only
communication is the batch split and the gather of variable-length results.  One process per GPU, full weight
replica per process; `torch.distributed` (NCCL on GPUs, gloo in the CPU tests) carries KB–MB messages only — no
collective sits next to a kernel, so there is nothing to fuse a collective into."""
from typing import Any, Callable, List, Optional, Sequence

import torch
import torch.distributed as dist


def shard_bounds(n_items: int, world: int, rank: int):
    """Contiguous balanced shard [lo, hi) of n_items for `rank` (first n%world ranks get one extra)."""
    q, r = divmod(n_items, world)
    lo = rank * q + min(rank, r)
    return lo, lo + q + (1 if rank < r else 0)


def length_balanced_order(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Greedy longest-first assignment of requests to ranks (total prompt length per rank as even as possible).
    Returns, per rank, the original indices it should run."""
    bins: List[List[int]] = [[] for _ in range(world)]
    load = [0] * world
    for i in sorted(range(len(lengths)), key=lambda k: -lengths[k]):
        r = min(range(world), key=lambda k: (load[k], k))
        bins[r].append(i)
        load[r] += lengths[i]
    return [sorted(b) for b in bins]


def run_data_parallel(fn: Callable[[List[Any]], List[Any]], requests: Sequence[Any], group: Optional[dist.ProcessGroup] = None,
                      lengths: Optional[Sequence[int]] = None) -> List[Any]:
    """Every rank calls this with the same `requests`; rank r runs `fn` on its shard only; every rank returns the
    full result list in request order.  Results must be picklable (numpy arrays / CPU tensors)."""
    if not dist.is_available() or not dist.is_initialized():
        return list(fn(list(requests)))
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    if lengths is not None:
        mine = length_balanced_order(lengths, world)[rank]
    else:
        lo, hi = shard_bounds(len(requests), world, rank)
        mine = list(range(lo, hi))
    local = list(fn([requests[i] for i in mine])) if mine else []
    assert len(local) == len(mine), "fn must return one result per request"
    gathered: List[Any] = [None] * world
    dist.all_gather_object(gathered, list(zip(mine, local)), group=group)
    out: List[Any] = [None] * len(requests)
    for part in gathered:
        for i, v in part:
            out[i] = v
    return out


def max_over_ranks(value: float, device=None) -> float:
    """Device-side MAX reduction of a timing (multi-GPU numbers are always the max over ranks)."""
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([value], dtype=torch.float64, device=device or ("cuda" if dist.get_backend() == "nccl" else "cpu"))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
