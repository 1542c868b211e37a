"""Host-side work partitioning over the GPUs of one box (one process per GPU).

SURVEY.md 8(e): the path shards by (1) unit-cell site -- the VUMPS eigenproblems of one iteration are
independent (the reference runs them as tasks, HubbardFunctions.jl:37) -- and (2) MPO virtual level --
y = sum_{a,b} GL[a] x W[a,b] GR[b] splits into per-rank partial sums over disjoint level sets followed
by one allreduce(sum) of y.  This module holds the pure bookkeeping (used by bench.py and tested with
world_size-2 gloo on CPU); the collective itself is torch.distributed plumbing.
"""
from __future__ import annotations


def site_for_rank(rank: int, unit_cell: int) -> int:
    """Site whose H_AC eigenproblem rank `rank` owns (replicas wrap round when world > unit cell)."""
    return rank % unit_cell


def aggregate_rate(world: int, steps: int, ms_per_rank) -> float:
    """Whole-job units/s: all ranks process `steps` units; time = max over ranks (device timed)."""
    return world * steps / (max(ms_per_rank) * 1e-3)


def partition_levels(level_pairs, chi: int, world: int):
    """Assign every MPO level pair (a, b) to a rank such that connected chains of levels (a term's whole
    level chain, the unit that keeps the W stage and the environment updates local) stay together and
    the ranks get a similar number of pairs.  Level 0 / chi-1 (the identity ends) do not glue chains.
    Returns owner[(a, b)] -> rank."""
    parent = list(range(chi))

    def find(i):
        while parent[i] != i:
            parent[i] = parent[parent[i]]
            i = parent[i]
        return i

    for (a, b) in level_pairs:
        if a in (0, chi - 1) or b in (0, chi - 1) or a == b:
            continue
        ra, rb = find(a), find(b)
        if ra != rb:
            parent[ra] = rb
    comp_pairs = {}
    for (a, b) in level_pairs:
        anchor = a if a not in (0, chi - 1) else (b if b not in (0, chi - 1) else None)
        key = ("ends",) if anchor is None else find(anchor)
        comp_pairs.setdefault(key, []).append((a, b))
    load = [0] * world
    owner = {}
    for key, pairs in sorted(comp_pairs.items(), key=lambda kv: -len(kv[1])):
        r = min(range(world), key=lambda i: load[i])
        load[r] += len(pairs)
        for p in pairs:
            owner[p] = r
    return owner


def partition_rows(level_pairs, chi: int, world: int):
    """Alternative for a single H_eff apply (no environment update involved): all pairs (a, .) of one left
    level a go to one rank, levels dealt out by descending pair count.  Stage L (GL[a] . x) then splits
    without duplication; stage R is duplicated where several ranks feed the same right level."""
    by_a = {}
    for (a, b) in level_pairs:
        by_a.setdefault(a, []).append((a, b))
    load = [0] * world
    owner = {}
    for a, pairs in sorted(by_a.items(), key=lambda kv: (-len(kv[1]), kv[0])):
        r = min(range(world), key=lambda i: load[i])
        load[r] += len(pairs)
        for p in pairs:
            owner[p] = r
    return owner


def shard_mpo_entries(entries: dict, chi: int, world: int, rank: int, mode: str = "chains") -> dict:
    """Entries {(a,s',s,b,c): w} of the MPO tensor owned by `rank`: the plan built from them computes this
    rank's partial sum of y = sum_{a,b} GL[a] x W[a,b] GR[b]; allreduce(sum) over the ranks gives y."""
    pairs = sorted({(k[0], k[3]) for k in entries})
    owner = partition_levels(pairs, chi, world) if mode == "chains" else partition_rows(pairs, chi, world)
    return {k: v for k, v in entries.items() if owner[(k[0], k[3])] == rank}
