#!/usr/bin/env python
"""Golden fixture for the batched stability check: inputs, the reference's own tables (as data) and
the outputs of the UNMODIFIED reference function `check_molecular_stability` (src/datamodules/components/edm/
__init__.py:91-124) per molecule, imported through oracle/ref_shim.py in the build container.
Run:  python tests/golden/make_golden_stability.py"""
import os
import sys

import numpy as np
import torch
import torch._dynamo  # noqa: F401  (before the stub modules are installed)

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import ref_shim  # noqa: E402

ref_shim.install()
from src.datamodules.components.edm import check_molecular_stability, get_bond_length_arrays, get_bond_order_batch  # noqa: E402
import src.datamodules.components.edm.constants as K  # noqa: E402
from src.datamodules.components.edm.datasets_config import QM9_WITH_H, GEOM_WITH_H  # noqa: E402


def molecule(rng, n, a, spacing):
    """Jittered lattice so that realistic bond lengths (and a few too-short / too-long pairs) occur."""
    side = int(np.ceil(n ** (1 / 3))) + 1
    grid = np.stack(np.meshgrid(*[np.arange(side)] * 3, indexing="ij"), -1).reshape(-1, 3)
    pick = rng.choice(len(grid), size=n, replace=False)
    pos = grid[pick] * spacing + rng.normal(0, 0.12, size=(n, 3))
    return pos.astype(np.float32), rng.integers(0, a, size=n)


def handmade(enc):
    """Methane, water, H2 and a stretched (broken) H2: stable / stable / stable / unstable."""
    t = 1.09 / np.sqrt(3.0)
    ch4 = (np.array([[0, 0, 0], [t, t, t], [t, -t, -t], [-t, t, -t], [-t, -t, t]], dtype=np.float32),
           np.array([enc["C"], enc["H"], enc["H"], enc["H"], enc["H"]]))
    h2o = (np.array([[0, 0, 0], [0.96, 0, 0], [-0.24, 0.93, 0]], dtype=np.float32), np.array([enc["O"], enc["H"], enc["H"]]))
    h2 = (np.array([[0, 0, 0], [0.74, 0, 0]], dtype=np.float32), np.array([enc["H"], enc["H"]]))
    h2x = (np.array([[0, 0, 0], [1.40, 0, 0]], dtype=np.float32), np.array([enc["H"], enc["H"]]))
    return [ch4, h2o, h2, h2x]


def case(info, sizes, seed, spacing, extra=False):
    rng = np.random.default_rng(seed)
    dec = list(info["atom_decoder"])
    enc = dict(info["atom_encoder"])
    b = get_bond_length_arrays(enc)
    di = dict(info)
    di["bonds1"], di["bonds2"], di["bonds3"] = b
    xs, ts, outs, es = [], [], [], []
    limit = "GEOM" in info["name"]
    mols = [molecule(rng, n, len(dec), spacing) for n in sizes] + (handmade(enc) if extra else [])
    sizes = [len(t) for _, t in mols]
    for p, t in mols:
        st, ns, nn = check_molecular_stability(torch.from_numpy(p), torch.from_numpy(np.asarray(t, dtype=np.int64)), di)
        xs.append(p); ts.append(np.asarray(t, dtype=np.int64)); outs.append((bool(st), int(ns), int(nn)))
        # the (A, E) graph make_mol_edm hands to RDKit (rdkit_functions.py:287-296; RDKit itself is not installed here):
        # the reference's own get_bond_order_batch on cartesian_prod(atom_types, atom_types), then tril(-1)
        pt, tt = torch.from_numpy(p), torch.from_numpy(np.asarray(t, dtype=np.int64))
        n = len(tt)
        dists = torch.cdist(pt.unsqueeze(0), pt.unsqueeze(0), p=2).squeeze(0).view(-1)
        a1, a2 = torch.cartesian_prod(tt, tt).T if n > 1 else (tt.repeat(1), tt.repeat(1))
        e_full = get_bond_order_batch(a1, a2, dists, di, limit_bonds_to_one=limit).view(n, n)
        es.append(torch.tril(e_full, diagonal=-1).to(torch.int8))
    return dict(atom_decoder=dec, bonds=[np.asarray(v, dtype=np.float32) for v in b],
                margins=(K.margin1, K.margin2, K.margin3), allowed_bonds={k: K.allowed_bonds[k] for k in dec},
                sizes=list(sizes), x=torch.from_numpy(np.concatenate(xs)), atom_types=torch.from_numpy(np.concatenate(ts)),
                ref=outs, bond_E=es, limit_bonds_to_one=limit)


fx = {"qm9": case(QM9_WITH_H, [19, 5, 23, 1, 12, 29, 2, 17], 3, 1.15, extra=True),
      "geom": case(GEOM_WITH_H, [44, 30, 61, 9, 25], 4, 1.3)}
torch.save(fx, os.path.join(os.path.dirname(os.path.abspath(__file__)), "stability.pt"))
for k, v in fx.items():
    print(k, v["ref"])
