"""2+ GPUs (torchrun, NCCL): the data-parallel training step of BASELINE config 5.  Every rank runs GCDMTrainLoss + backward on
its OWN batch, the gradients are averaged with bdiff.distributed.allreduce_mean_ (one bucketed NCCL all-reduce) and the
optimiser kernels step.  Checks: (1) the averaged gradient equals the mean of the per-rank gradients recomputed serially on
rank 0's GPU (bit-exact inputs, fp32 sum order differs: <= 1e-6 of the norm); (2) after three steps all ranks hold identical
parameters and EMA weights (bit-exact)."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "bio-diffusion_b200"), os.path.join(ROOT, "oracle")):
    sys.path.insert(0, p)
import bdiff  # noqa: E402
import gcpnet_oracle as O  # noqa: E402  (seeded weights only)
from bdiff.datasets import GEOM_N_NODES, sample_num_nodes  # noqa: E402
from bdiff.distributed import allreduce_mean_  # noqa: E402
from bdiff.optim import GCDMTrainTail  # noqa: E402


def make_batch(r, b, dev, B=8):
    sizes = sample_num_nodes(GEOM_N_NODES, B, seed=100 * r + b)
    g = torch.Generator().manual_seed(7 * r + b)
    bi = torch.repeat_interleave(torch.arange(B), sizes)
    n = bi.shape[0]
    x = torch.randn((n, 3), generator=g) * 2
    x = x - (torch.zeros((B, 3)).index_add_(0, bi, x) / sizes[:, None].float())[bi]
    oh = torch.nn.functional.one_hot(torch.randint(0, 16, (n,), generator=g), 16).float()
    t_int = torch.randint(0, 1001, (B, 1), generator=g)
    noise = [torch.randn((n, 3), generator=g), torch.randn((n, 16), generator=g)]
    return tuple(v.to(dev) for v in (bi, torch.ones(n, dtype=torch.bool), x, oh, torch.zeros((n, 0)))), t_int, noise


def loss_of(tl, batch, t_int, noise, dev):
    it = iter(noise)
    return tl(*batch, None, t_int=t_int, noise=lambda s: next(it).to(dev))[0].mean()


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    sd = O.random_state_dict(O.config_named("geom"), seed=7, scale=0.7)
    net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named("geom"))
    net.load_state_dict(sd, strict=True)
    net.to(dev)
    net.flatten_parameters()
    opt = GCDMTrainTail(net.parameters(), lr=1e-4)
    tl = bdiff.GCDMTrainLoss(net, GEOM_N_NODES)
    # (1) averaged gradient
    opt.zero_grad()
    loss_of(tl, *make_batch(rank, 0, dev), dev).backward()
    n_coll = allreduce_mean_(opt.grads)
    avg = torch.cat([g.reshape(-1) for g in opt.grads]).clone()
    if rank == 0:
        acc = torch.zeros_like(avg)
        for r in range(world):
            opt.zero_grad()
            loss_of(tl, *make_batch(r, 0, dev), dev).backward()
            acc += torch.cat([g.reshape(-1) for g in opt.grads])
        acc /= world
        rel = ((avg - acc).norm() / acc.norm()).item()
        print(f"ddp_train_check: world {world}, {n_coll} all-reduce per step, averaged gradient vs serial mean: rel {rel:.2e}")
        assert rel < 1e-6, rel
    dist.barrier()
    # (2) three steps, identical replicas afterwards
    for b in range(1, 4):
        opt.zero_grad()
        loss_of(tl, *make_batch(rank, b, dev), dev).backward()
        allreduce_mean_(opt.grads)
        opt.step()
    flat = torch.cat([p.detach().reshape(-1) for p in net.parameters()] + [e.reshape(-1) for e in opt.ema_parameters()])
    ref = flat.clone()
    dist.broadcast(ref, 0)
    same = torch.tensor([int(torch.equal(flat, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    if rank == 0:
        print(f"ddp_train_check: replicas identical after 3 steps: {bool(same.item())}; report {opt.report()['step']} steps")
    assert bool(same.item())
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
