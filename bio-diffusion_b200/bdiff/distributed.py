"""Multi-GPU sampling: shard independent molecules across ranks, one final gather.

The reference samples on one device only (src/mol_gen_sample.py:108-112).  Molecules are independent except
for the `_orientations` boundary quirk, so each rank runs the whole chain on its own
sub-batch with no communication and the final [N_r, 3+A(+1)] blocks are exchanged once (NCCL all_gather of
padded blocks over NVLink; gloo on CPU for tests).  Parity policy: PER-SHARD — the oracle for rank r is the
reference run on rank r's sub-batch (what a user sharding the reference by hand would get).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def lpt_shards(num_nodes: Sequence[int], world_size: int) -> List[List[int]]:
    """Longest-processing-time bin packing of molecules by cost n^2 (edge count). Deterministic.

    Returns, per rank, the molecule ids it owns (ascending, so each shard keeps the caller's order).
    """
    order = sorted(range(len(num_nodes)), key=lambda i: (-int(num_nodes[i]) ** 2, i))
    loads = [0] * world_size
    shards: List[List[int]] = [[] for _ in range(world_size)]
    for i in order:
        r = min(range(world_size), key=lambda k: (loads[k], k))
        shards[r].append(i)
        loads[r] += int(num_nodes[i]) ** 2
    return [sorted(s) for s in shards]


def gather_results(local_out: torch.Tensor, local_mols: Sequence[int], num_nodes: Sequence[int],
                   world_size: int, group=None) -> torch.Tensor:
    """all_gather the per-rank [N_r, D] blocks and restore the global molecule order -> [N, D] on every rank."""
    shards = lpt_shards(num_nodes, world_size)
    counts = [sum(int(num_nodes[i]) for i in s) for s in shards]
    d = local_out.shape[1]
    pad = max(max(counts), 1)
    buf = torch.zeros((pad, d), dtype=local_out.dtype, device=local_out.device)
    buf[: local_out.shape[0]] = local_out
    if world_size > 1:
        gathered = [torch.empty_like(buf) for _ in range(world_size)]
        dist.all_gather(gathered, buf, group=group)
    else:
        gathered = [buf]
    offsets = [0]
    for n in num_nodes:
        offsets.append(offsets[-1] + int(n))
    out = torch.empty((offsets[-1], d), dtype=local_out.dtype, device=local_out.device)
    for r, s in enumerate(shards):
        pos = 0
        for i in s:
            n = int(num_nodes[i])
            out[offsets[i]: offsets[i] + n] = gathered[r][pos: pos + n]
            pos += n
    return out


def sample_sharded(sampler, num_nodes: torch.Tensor, context: Optional[torch.Tensor] = None,
                   num_timesteps: Optional[int] = None, group=None, gather: bool = True):
    """Each rank samples its LPT shard with `sampler` (a GCDMSampler); returns (out_full or out_local, my_mols)."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    sizes = [int(v) for v in num_nodes.tolist()]
    mine = lpt_shards(sizes, world)[rank]
    idx = torch.tensor(mine, dtype=torch.long)
    local_nodes = num_nodes.cpu()[idx]
    local_ctx = context.cpu()[idx] if context is not None else None
    out, _, _ = sampler.sample(local_nodes, local_ctx, num_timesteps)
    if not gather:
        return out, mine
    return gather_results(out, mine, sizes, world, group), mine
