"""CPU restatement of the molecular-stability check.  TEST INFRASTRUCTURE ONLY.

Follows /root/reference/src/datamodules/components/edm/__init__.py:
  * `get_bond_order_batch`        :61-88   (distances * 100, three table lookups + margins, later assignments overwrite)
  * `check_molecular_stability`   :91-124  (pairwise distances, diagonal zeroed, row sums, `allowed_bonds` membership;
                                            returns (molecule_stable, nr_stable_bonds, n))
with the tables of edm/constants.py (`bonds1/2/3`, `margin1/2/3` = 10/5/3 pm, `allowed_bonds`) passed in as DATA.
One deliberate difference: distances are the direct fp32 expression sqrt((dx*dx + dy*dy) + dz*dz) for every n, while
`torch.cdist` switches to a matmul formulation above 25 atoms (same value up to fp32 round-off, which can only matter for
a pair sitting exactly on a threshold).  Pinned against the reference's own function on the fixtures of
tests/golden/make_golden_stability.py (identical integers on all of them, including 44-atom molecules).
"""
import numpy as np


def allowed_mask(atom_decoder, allowed_bonds):
    """bit c of mask[type] set <=> an atom of that type may have c bonds (constants.py `allowed_bonds`: int or list)."""
    out = np.zeros(len(atom_decoder), dtype=np.uint32)
    for i, sym in enumerate(atom_decoder):
        v = allowed_bonds[sym]
        for c in ([v] if isinstance(v, int) else list(v)):
            out[i] |= np.uint32(1) << np.uint32(c)
    return out


def check_stability_batch(x, atom_types, mol_off, bonds, margins, mask, limit_bonds_to_one=False):
    """x [N,3] fp32, atom_types [N] int, mol_off [B+1]; bonds = (b1, b2, b3) [A,A]; returns nr_bonds [N], nr_stable [B],
    mol_stable [B] (all integer arrays)."""
    x = np.asarray(x, dtype=np.float32)
    t = np.asarray(atom_types, dtype=np.int64)
    b1, b2, b3 = (np.asarray(b, dtype=np.float32) for b in bonds)
    m1, m2, m3 = (np.float32(m) for m in margins)
    nb = np.zeros(len(t), dtype=np.int32)
    nr_stable = np.zeros(len(mol_off) - 1, dtype=np.int32)
    mol_stable = np.zeros(len(mol_off) - 1, dtype=np.int32)
    for k in range(len(mol_off) - 1):
        a, b = int(mol_off[k]), int(mol_off[k + 1])
        p, tt = x[a:b], t[a:b]
        d = p[:, None, :] - p[None, :, :]                                    # [n,n,3] fp32
        d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
        dist = np.float32(100.0) * np.sqrt(d2)                               # :69 "we change the metric"
        idx = (tt[:, None], tt[None, :])
        order = np.zeros(dist.shape, dtype=np.int32)
        order[dist < b1[idx] + m1] = 1                                       # :76-81
        order[dist < b2[idx] + m2] = 2
        order[dist < b3[idx] + m3] = 3
        if limit_bonds_to_one:
            order[order > 1] = 1
        np.fill_diagonal(order, 0)                                           # :111
        s = order.sum(axis=1)
        nb[a:b] = s
        ok = [(int(c) < 32 and (int(mask[ti]) >> int(c)) & 1) for ti, c in zip(tt, s)]   # :114-121
        nr_stable[k] = int(sum(ok))
        mol_stable[k] = int(nr_stable[k] == b - a)
    return nb, nr_stable, mol_stable


def bond_order_matrix(x, atom_types, bonds, margins, limit_bonds_to_one=False):
    """E of `make_mol_edm` (rdkit_functions.py:287-296) for ONE molecule: tril(get_bond_order_batch(type_i, type_j, dist), -1)
    as an int [n, n] array (atoms1, atoms2 = cartesian_prod(atom_types, atom_types).T -> (type_i, type_j) at i*n + j)."""
    p = np.asarray(x, dtype=np.float32)
    tt = np.asarray(atom_types, dtype=np.int64)
    b1, b2, b3 = (np.asarray(b, dtype=np.float32) for b in bonds)
    m1, m2, m3 = (np.float32(m) for m in margins)
    d = p[:, None, :] - p[None, :, :]
    d2 = (d[..., 0] * d[..., 0] + d[..., 1] * d[..., 1]) + d[..., 2] * d[..., 2]
    dist = np.float32(100.0) * np.sqrt(d2)
    idx = (tt[:, None], tt[None, :])
    order = np.zeros(dist.shape, dtype=np.int64)
    order[dist < b1[idx] + m1] = 1
    order[dist < b2[idx] + m2] = 2
    order[dist < b3[idx] + m3] = 3
    if limit_bonds_to_one:
        order[order > 1] = 1
    return np.tril(order, -1)
