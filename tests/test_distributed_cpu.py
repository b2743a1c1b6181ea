"""CPU, world_size 2 over gloo: molecule sharding (LPT by n^2) + the single final gather restore the global
order — the N>1 path of bench.py / bdiff.distributed without a GPU (the per-rank sampler is faked)."""
import os
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


class FakeSampler:
    """Stands in for GCDMSampler: row j of molecule with n atoms is [n, j, 100 n + j]."""
    from types import SimpleNamespace
    cfg = SimpleNamespace(num_atom_types=0, include_charges=False)      # output width 3

    def _device(self):
        return torch.device("cpu")

    def sample(self, num_nodes, context=None, num_timesteps=None):
        rows = []
        for n in num_nodes.tolist():
            for j in range(n):
                rows.append([float(n), float(j), float(100 * n + j)])
        out = torch.tensor(rows).reshape(-1, 3)
        return out, None, None


def _worker(rank, world, port, sizes, results):
    sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bdiff.distributed import sample_sharded
    out, mine = sample_sharded(FakeSampler(), torch.tensor(sizes))
    results[rank] = (out, mine)
    dist.barrier()
    dist.destroy_process_group()


def test_lpt_shards_balance_and_cover():
    sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
    from bdiff.distributed import lpt_shards
    sizes = [181, 3, 44, 44, 30, 61, 25, 19, 19, 100, 7, 90]
    shards = lpt_shards(sizes, 4)
    assert sorted(i for s in shards for i in s) == list(range(len(sizes)))
    loads = [sum(sizes[i] ** 2 for i in s) for s in shards]
    assert max(loads) <= 181 ** 2 + 1      # the biggest molecule bounds the best possible makespan here
    assert lpt_shards(sizes, 1) == [list(range(len(sizes)))]
    assert lpt_shards([19] * 8, 4) == [[0, 4], [1, 5], [2, 6], [3, 7]]


def test_two_rank_gather_restores_global_order():
    sizes = [5, 19, 3, 12, 7, 19, 2]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, 29517, sizes, results), nprocs=2, join=True)
    expect, _, _ = FakeSampler().sample(torch.tensor(sizes))
    for rank in (0, 1):
        out, mine = results[rank]
        assert torch.equal(out, expect)
    assert sorted(results[0][1] + results[1][1]) == list(range(len(sizes)))


def _grad_worker(rank, world, port, results):
    sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from bdiff.distributed import allreduce_mean_, allreduce_mean_flat_
    g = torch.Generator().manual_seed(100 + rank)
    shapes = [(256, 605), (256,), (32, 8), (1, 256), (17,)]
    grads = [torch.randn(s, generator=g) for s in shapes]
    mine = [t.clone() for t in grads]
    n_one = allreduce_mean_(grads)                       # everything in one bucket
    again = [t.clone() for t in mine]
    n_many = allreduce_mean_(again, bucket_bytes=4096)   # forced into several buckets: same result
    flat = torch.cat([t.reshape(-1) for t in mine])       # the optimiser tail's layout: views of one buffer
    views, o = [], 0
    for t in mine:
        views.append(flat[o:o + t.numel()].view_as(t))
        o += t.numel()
    n_flat = allreduce_mean_flat_(flat)
    results[rank] = (mine, grads, again, n_one, n_many, [v.clone() for v in views], n_flat)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gradient_allreduce_mean():
    """Config 5's only collective: the mean of every rank's gradients, one all-reduce for the whole model."""
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_grad_worker, args=(2, 29533, results), nprocs=2, join=True)
    (a0, r0, m0, n_one0, n_many0, f0, nf0), (a1, r1, m1, n_one1, n_many1, f1, nf1) = results[0], results[1]
    assert n_one0 == n_one1 == 1 and n_many0 == n_many1 and n_many0 > 1 and nf0 == nf1 == 1
    for x0, x1, y0, y1, z0, v0, v1 in zip(a0, a1, r0, r1, m0, f0, f1):
        want = (x0 + x1) / 2
        assert torch.allclose(y0, want, rtol=0, atol=1e-7) and torch.equal(y0, y1)
        assert torch.allclose(z0, want, rtol=0, atol=1e-7)
        assert torch.equal(v0, y0) and torch.equal(v0, v1)          # flat-buffer exchange == bucketed exchange


def test_allreduce_mean_single_process_is_noop():
    sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
    from bdiff.distributed import allreduce_mean_
    t = [torch.ones(3)]
    assert allreduce_mean_(t) == 0 and torch.equal(t[0], torch.ones(3))


def test_more_ranks_than_molecules_does_not_hang():
    """a rank with an empty shard must still take part in the final gather."""
    sizes = [6]
    mgr = mp.Manager()
    results = mgr.dict()
    mp.spawn(_worker, args=(2, 29519, sizes, results), nprocs=2, join=True)
    expect, _, _ = FakeSampler().sample(torch.tensor(sizes))
    for rank in (0, 1):
        assert torch.equal(results[rank][0], expect)
    assert results[1][1] == []


def test_shard_imbalance_metric():
    sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
    from bdiff.distributed import shard_imbalance
    assert shard_imbalance([19] * 8, 4) == 1.0
    assert shard_imbalance([181, 3, 3, 3], 2) > 1.9
