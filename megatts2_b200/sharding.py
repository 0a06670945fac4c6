"""Multi-GPU batch split: utterances are independent, so the path shards by
contiguous batch slices with a full weight replica per GPU and NO data-path collective.  One
process per GPU (torchrun); torch.distributed (NCCL over NVLink on GPU, gloo in the CPU tests) is
used only to gather per-utterance results / lengths at the end, and for the bench barrier."""
from typing import List, Sequence, Tuple

import torch


def shard_bounds(n_items: int, world: int) -> List[Tuple[int, int]]:
    """Contiguous, balanced [lo, hi) slices; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    out, lo = [], 0
    for r in range(world):
        hi = lo + base + (1 if r < extra else 0)
        out.append((lo, hi))
        lo = hi
    return out


def my_shard(n_items: int, rank: int, world: int) -> Tuple[int, int]:
    return shard_bounds(n_items, world)[rank]


def balance_by_cost(costs: Sequence[float], world: int) -> List[List[int]]:
    """Greedy longest-processing-time assignment of utterance indices to ranks.  The AR decode cost of an
    utterance grows ~quadratically with its length, so length bucketing across ranks is the only
    load-balance concern of the split."""
    order = sorted(range(len(costs)), key=lambda i: -costs[i])
    loads = [0.0] * world
    bins: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min(range(world), key=lambda j: loads[j])
        bins[r].append(i)
        loads[r] += costs[i]
    return [sorted(b) for b in bins]


def gather_variable(local: torch.Tensor, lengths: torch.Tensor, group=None):
    """All-gather per-utterance rows of different lengths: local (b_local, Lmax_local) + lengths (b_local,)
    -> list over ranks of (tensor, lengths).  Pads to the global max length for the collective."""
    import torch.distributed as dist
    world = dist.get_world_size(group)
    meta = torch.tensor([local.shape[0], local.shape[1]], dtype=torch.int64, device=local.device)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    bmax = int(max(m[0] for m in metas))
    lmax = int(max(m[1] for m in metas))
    buf = torch.zeros(bmax, lmax, dtype=local.dtype, device=local.device)
    buf[: local.shape[0], : local.shape[1]] = local
    lens = torch.zeros(bmax, dtype=torch.int64, device=local.device)
    lens[: lengths.shape[0]] = lengths.to(torch.int64)
    bufs = [torch.zeros_like(buf) for _ in range(world)]
    lenss = [torch.zeros_like(lens) for _ in range(world)]
    dist.all_gather(bufs, buf, group=group)
    dist.all_gather(lenss, lens, group=group)
    return [(bufs[r][: int(metas[r][0]), : int(metas[r][1])], lenss[r][: int(metas[r][0])]) for r in range(world)]
