#!/usr/bin/env python
"""Golden fixture for the packed collation: a synthetic padded QM9-style dataset, a batch of molecule ids,
and what the UNMODIFIED reference makes of it — `ProcessedDataset._featurize_as_graph` per molecule
(src/datamodules/components/edm_dataset.py:187-216), concatenated like PyG's collater does, and `prepare_context`
(src/datamodules/components/edm/utils.py:333-382) — imported through oracle/ref_shim.py in the build container.
Run:  python tests/golden/make_golden_collate.py"""
import os
import sys

import torch
import torch._dynamo  # noqa: F401

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref_shim.install()
from src.datamodules.components.edm_dataset import ProcessedDataset  # noqa: E402
from src.datamodules.components.edm.utils import prepare_context  # noqa: E402
from torch_geometric.data import Batch  # noqa: E402  (the shim's attribute bag)

g = torch.Generator().manual_seed(17)
M, P = 40, 29
species = torch.tensor([1, 6, 7, 8, 9])
natoms = torch.randint(3, P + 1, (M,), generator=g)
natoms[5] = P                      # a full molecule
natoms[9] = 1                      # a single atom
charges = torch.zeros((M, P), dtype=torch.int64)
for m in range(M):
    charges[m, : natoms[m]] = species[torch.randint(0, 5, (int(natoms[m]),), generator=g)]
positions = torch.randn((M, P, 3), generator=g)                   # padding rows hold garbage on purpose
one_hot = charges.unsqueeze(-1) == species.view(1, 1, -1)
alpha = torch.randn(M, generator=g) * 8 + 75
mu = torch.randn(M, generator=g)
idx = torch.tensor([7, 5, 9, 0, 33, 12, 21, 5])                   # a repeated id is legal

graphs = []
for m in idx.tolist():
    mol = {"index": torch.tensor(m), "positions": positions[m].clone(), "charges": charges[m].clone(),
           "one_hot": one_hot[m].clone(), "num_atoms": natoms[m], "alpha": alpha[m], "mu": mu[m]}
    graphs.append(ProcessedDataset._featurize_as_graph(None, mol))
# PyG collation = concatenation of the per-graph tensors + the `batch` vector
cat = lambda k: torch.cat([getattr(d, k) for d in graphs], 0)     # noqa: E731
batch = Batch(x=cat("x"), one_hot=cat("one_hot"), charges=cat("charges"), mask=cat("mask"),
              batch=torch.repeat_interleave(torch.arange(len(graphs)), P), index=idx.unsqueeze(-1),
              alpha=torch.stack([d.alpha.reshape(()) for d in graphs]), mu=torch.stack([d.mu.reshape(()) for d in graphs]))
norms = {"alpha": {"mean": alpha.mean(), "mad": (alpha - alpha.mean()).abs().mean()},
         "mu": {"mean": mu.mean(), "mad": (mu - mu.mean()).abs().mean()}}
ctx = prepare_context(["alpha", "mu"], batch, norms)
keep = batch.mask
fx = dict(positions=positions, charges=charges, one_hot=one_hot, alpha=alpha, mu=mu, idx=idx, norms=norms, pad=P,
          ref=dict(x=batch.x[keep], one_hot=batch.one_hot[keep], charges=batch.charges[keep].unsqueeze(-1),
                   batch=batch.batch[keep], context=ctx[keep], padded_rows=int(keep.numel()), present_rows=int(keep.sum())))
torch.save(fx, os.path.join(os.path.dirname(os.path.abspath(__file__)), "collate.pt"))
print("padded rows", fx["ref"]["padded_rows"], "present rows", fx["ref"]["present_rows"], "context", tuple(ctx.shape))
