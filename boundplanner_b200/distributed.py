"""Multi-GPU sharding of the set-graph build (one process per GPU).

The path shards naturally: seeds are independent given the
replicated obstacle table, set pairs are independent given all sets.  The only
exchange step is an all-gather of the halfspace tensors (NCCL over NVLink on
the GPU box, gloo in the CPU tests), followed by a row-block partition of the
S x S pair matrix and an all-gather of the adjacency bit-rows.

Everything here is host-side orchestration; it works on any torch device so
the partition / gather logic is covered by world_size-2 gloo tests on CPU, with
the pair kernel injected as a callable.
"""
from __future__ import annotations

import torch
import torch.distributed as dist


def shard_range(n, rank, world):
    """Contiguous block of ``n`` items owned by ``rank`` (first ranks get the remainder)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def balanced_row_blocks(S, world):
    """Row ranges [r0,r1) of the strict upper triangle with (nearly) equal pair
    counts: row i holds S-1-i pairs, so early rows are heavier."""
    total = S * (S - 1) // 2
    bounds = [0]
    acc = 0
    target = 1
    for i in range(S):
        acc += S - 1 - i
        while target < world and acc >= total * target / world:
            bounds.append(i + 1)
            target += 1
    while len(bounds) < world:
        bounds.append(S)
    bounds.append(S)
    return [(bounds[r], bounds[r + 1]) for r in range(world)]


def all_gather_sets(A, b, m, group=None):
    """All-gather the local halfspace tensors [S_loc,m_max,3], [S_loc,m_max], [S_loc]
    (equal S_loc on every rank) into the global ones, rank-major."""
    world = dist.get_world_size(group)
    if world == 1:
        return A, b, m
    # one packed message: rows (a0,a1,a2,b) + the row count as a double in an extra row
    S_loc, m_max = b.shape
    packed = torch.empty((S_loc, m_max + 1, 4), dtype=A.dtype, device=A.device)
    packed[:, :m_max, :3] = A
    packed[:, :m_max, 3] = b
    packed[:, m_max, :] = m.to(A.dtype).unsqueeze(1)
    out = torch.empty((world * S_loc, m_max + 1, 4), dtype=A.dtype, device=A.device)
    dist.all_gather_into_tensor(out, packed, group=group)
    Ag = out[:, :m_max, :3].contiguous()
    bg = out[:, :m_max, 3].contiguous()
    mg = out[:, m_max, 0].to(torch.int32).contiguous()
    return Ag, bg, mg


def sharded_adjacency(A, b, m, pair_fn, tol=0.01, group=None):
    """Global adjacency bit-matrix [S, words] (int32 words) from local sets.

    pair_fn(A, b, m, tol, row_begin, row_end) -> int32 [rows, words] is the pair
    kernel (boundplanner_b200.geometry.pair_feasible on the GPU)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    if world == 1:
        return pair_fn(A, b, m, tol, 0, A.shape[0]), (A, b, m)
    Ag, bg, mg = all_gather_sets(A, b, m, group)
    S = Ag.shape[0]
    words = (S + 31) // 32
    blocks = balanced_row_blocks(S, world)
    r0, r1 = blocks[rank]
    local = pair_fn(Ag, bg, mg, tol, r0, r1)
    # variable-size row blocks: pad to the largest block for a single all_gather
    max_rows = max(hi - lo for lo, hi in blocks)
    padded = torch.zeros((max_rows, words), dtype=torch.int32, device=Ag.device)
    padded[: r1 - r0] = local
    gathered = torch.empty((world * max_rows, words), dtype=torch.int32, device=Ag.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    bits = torch.cat([gathered[r * max_rows: r * max_rows + (hi - lo)] for r, (lo, hi) in enumerate(blocks)])
    return bits, (Ag, bg, mg)
