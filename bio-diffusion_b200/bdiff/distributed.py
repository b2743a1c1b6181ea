"""Multi-GPU sampling: shard independent molecules across ranks, one final gather.

The reference samples on one device only (src/mol_gen_sample.py:108-112).  Molecules are independent except
for the `_orientations` boundary quirk, so each rank runs the whole chain on its own
sub-batch with no communication and the final [N_r, 3+A(+1)] blocks are exchanged once (NCCL all_gather of
padded blocks over NVLink; gloo on CPU for tests).  Parity policy: PER-SHARD — the oracle for rank r is the
reference run on rank r's sub-batch (what a user sharding the reference by hand would get).

Training (config 5) is plain replica data parallelism in the reference (Lightning DDP, one all-reduce of all gradients
per step, configs/trainer/ddp.yaml); `allreduce_mean_` below is that exchange for a list of gradient tensors: packed
into buckets, ONE collective per bucket, averaged, unpacked in place.  
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


_rank_seed_folded = False


def decorrelate_rank_rng(rank: int) -> None:
    """Fold the rank into this process's default CUDA generator ONCE.  The reference seeds every process identically
    (`seed_everything(cfg.seed)`); with LPT giving equal-shaped shards for uniform-size batches, identically seeded
    ranks would draw the same noise and generate bit-identical molecules (silent duplicates in uniqueness / novelty
    statistics).  Rank 0 keeps the caller's stream, so a 1-GPU run is unchanged."""
    global _rank_seed_folded
    if _rank_seed_folded or rank == 0 or not torch.cuda.is_available():
        _rank_seed_folded = True
        return
    torch.cuda.manual_seed((torch.cuda.initial_seed() + 0x9E3779B97F4A7C15 * rank) % (1 << 63))
    _rank_seed_folded = True


def sample_sharded(sampler, num_nodes: torch.Tensor, context: Optional[torch.Tensor] = None,
                   num_timesteps: Optional[int] = None, group=None, gather: bool = True):
    """Each rank samples its LPT shard with `sampler` (a GCDMSampler); returns (out_full or out_local, my_mols).

    The per-rank noise streams are decorrelated here (see decorrelate_rank_rng).  A rank whose shard is empty (more
    ranks than molecules) skips the chain and contributes a zero-row block, so the final collective still matches."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    decorrelate_rank_rng(rank)
    sizes = [int(v) for v in num_nodes.tolist()]
    mine = lpt_shards(sizes, world)[rank]
    if mine:
        idx = torch.tensor(mine, dtype=torch.long)
        local_nodes = num_nodes.cpu()[idx]
        local_ctx = context.cpu()[idx] if context is not None else None
        out, _, _ = sampler.sample(local_nodes, local_ctx, num_timesteps)
    else:
        cfg = sampler.cfg
        out = torch.zeros((0, 3 + cfg.num_atom_types + int(cfg.include_charges)), device=sampler._device())
    if not gather:
        return out, mine
    return gather_results(out, mine, sizes, world, group), mine


def shard_imbalance(num_nodes: Sequence[int], world_size: int) -> float:
    """max over ranks of the LPT shard cost (sum n^2) divided by the mean: 1.0 = perfectly balanced."""
    shards = lpt_shards(num_nodes, world_size)
    loads = [sum(int(num_nodes[i]) ** 2 for i in s) for s in shards]
    mean = sum(loads) / max(1, world_size)
    return max(loads) / mean if mean > 0 else 1.0


def allreduce_mean_flat_(flat: torch.Tensor, group=None) -> int:
    """In-place mean over the ranks of ONE contiguous buffer (GCDMTrainTail keeps all gradients in one): a single
    all-reduce, no packing.  Returns the number of collectives issued (0 when the world size is 1)."""
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    flat.div_(dist.get_world_size(group))
    return 1


def allreduce_mean_(tensors: Sequence[torch.Tensor], group=None, bucket_bytes: int = 64 << 20) -> int:
    """In-place mean over the ranks of `group` of every tensor in `tensors` (the gradients of one step), packed into
    contiguous buckets of at most `bucket_bytes` so that a 2.7 M-parameter model is ONE all-reduce.  All tensors must
    share dtype and device.  Returns the number of collectives issued (0 when the world size is 1)."""
    tensors = [t for t in tensors if t is not None and t.numel() > 0]
    if not tensors:
        return 0
    if not dist.is_available() or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return 0
    world = dist.get_world_size(group)
    dtype, dev = tensors[0].dtype, tensors[0].device
    if any(t.dtype != dtype or t.device != dev for t in tensors):
        raise ValueError("allreduce_mean_: tensors must share dtype and device")
    per = max(1, bucket_bytes // tensors[0].element_size())
    buckets: List[List[torch.Tensor]] = [[]]
    fill = 0
    for t in tensors:
        if buckets[-1] and fill + t.numel() > per:
            buckets.append([])
            fill = 0
        buckets[-1].append(t)
        fill += t.numel()
    for b in buckets:
        flat = torch.cat([t.reshape(-1) for t in b])
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
        o = 0
        for t in b:
            t.copy_(flat[o:o + t.numel()].view_as(t))
            o += t.numel()
    return len(buckets)
