"""Token-balanced sharding of a ragged global batch over data-parallel ranks (SURVEY.md §7 "straggler risk", §8e).

The reference shards images over ranks with a plain DistributedSampler; with 1..10 channels per image the work per rank then
varies by sigma ~ 6.5 % (tokens) and every step runs at the pace of the most loaded rank.  Here the global batch's channel
counts (host ints, known at collate time: src/data/channels_strategies.py:31-85) are dealt longest-processing-time-first to
the ranks, under the constraint that every rank receives the same number of images.  Which images form the GLOBAL batch is
unchanged, so the optimisation problem is the reference's; only the assignment of images to replicas differs.
"""
from __future__ import annotations

from typing import List, Sequence

# cost model of one image with C channels (packed tokens S = 1 + C*N): linear layers ~ S, attention ~ S^2
_LINEAR_FLOPS_PER_TOKEN = 12 * 1_867_776.0
_ATTN_FLOPS_PER_S2 = 12 * 4 * 192.0


def image_cost(channels: int, npatch: int = 196) -> float:
    s = 1.0 + channels * npatch
    return _LINEAR_FLOPS_PER_TOKEN * s + _ATTN_FLOPS_PER_S2 * s * s


def token_balanced_shards(counts: Sequence[int], world: int, npatch: int = 196) -> List[List[int]]:
    """Indices of ``counts`` assigned to each of ``world`` ranks (equal cardinality; len(counts) must divide by world)."""
    n = len(counts)
    if world < 1 or n % world:
        raise ValueError(f"global batch of {n} images does not split evenly over {world} ranks")
    per = n // world
    order = sorted(range(n), key=lambda i: (-image_cost(int(counts[i]), npatch), i))
    load = [0.0] * world
    shards: List[List[int]] = [[] for _ in range(world)]
    for i in order:
        r = min((r for r in range(world) if len(shards[r]) < per), key=lambda r: (load[r], r))
        shards[r].append(i)
        load[r] += image_cost(int(counts[i]), npatch)
    return [sorted(s) for s in shards]
