"""GPU: bdiff_check_stability (one kernel for the whole batch) against the oracle and the stored outputs of the
reference's check_molecular_stability — integers, bit-exact."""
import os

import numpy as np
import pytest
import torch

import stability_oracle as SO
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", ["qm9", "geom"])
def test_batched_stability_matches_reference_and_oracle(name):
    from bdiff.stability import check_molecular_stability_batch
    fx = torch.load(os.path.join(GOLDEN, "stability.pt"), weights_only=False)[name]
    info = {"atom_decoder": fx["atom_decoder"], "bonds1": fx["bonds"][0], "bonds2": fx["bonds"][1], "bonds3": fx["bonds"][2]}
    stable, nr_stable, n, nr_bonds = check_molecular_stability_batch(
        fx["x"].cuda(), fx["atom_types"].cuda(), torch.tensor(fx["sizes"]), info, fx["allowed_bonds"], fx["margins"])
    got = list(zip(stable.cpu().tolist(), nr_stable.cpu().tolist(), n.cpu().tolist()))
    assert got == [tuple(r) for r in fx["ref"]]
    off = np.concatenate(([0], np.cumsum(fx["sizes"])))
    nb, ns, ms = SO.check_stability_batch(fx["x"].numpy(), fx["atom_types"].numpy(), off, fx["bonds"], fx["margins"],
                                          SO.allowed_mask(fx["atom_decoder"], fx["allowed_bonds"]))
    assert np.array_equal(nr_bonds.cpu().numpy(), nb)


def test_batched_stability_large_random_batch_vs_oracle():
    """512 molecules of 3..60 atoms (GEOM decoder): every per-atom bond count identical to the oracle."""
    from bdiff.stability import check_molecular_stability_batch
    fx = torch.load(os.path.join(GOLDEN, "stability.pt"), weights_only=False)["geom"]
    rng = np.random.default_rng(0)
    sizes = rng.integers(3, 61, size=512)
    n = int(sizes.sum())
    x = (rng.normal(0, 1.6, size=(n, 3))).astype(np.float32)
    t = rng.integers(0, len(fx["atom_decoder"]), size=n)
    info = {"atom_decoder": fx["atom_decoder"], "bonds1": fx["bonds"][0], "bonds2": fx["bonds"][1], "bonds3": fx["bonds"][2]}
    stable, nr_stable, nn, nr_bonds = check_molecular_stability_batch(
        torch.from_numpy(x).cuda(), torch.from_numpy(t).cuda(), torch.from_numpy(sizes), info, fx["allowed_bonds"], fx["margins"])
    off = np.concatenate(([0], np.cumsum(sizes)))
    nb, ns, ms = SO.check_stability_batch(x, t, off, fx["bonds"], fx["margins"],
                                          SO.allowed_mask(fx["atom_decoder"], fx["allowed_bonds"]))
    assert np.array_equal(nr_bonds.cpu().numpy(), nb)
    assert np.array_equal(nr_stable.cpu().numpy(), ns) and np.array_equal(stable.cpu().numpy().astype(np.int32), ms)
    assert nb.max() > 0


@pytest.mark.parametrize("name", ["qm9", "geom"])
def test_bond_order_matrix_matches_reference_graph(name):
    """bdiff_bond_orders vs the (A, E) graph of the reference's make_mol_edm (stored fixture): same bonds, same order."""
    from bdiff.stability import bond_orders_batch
    fx = torch.load(os.path.join(GOLDEN, "stability.pt"), weights_only=False)[name]
    info = {"atom_decoder": fx["atom_decoder"], "bonds1": fx["bonds"][0], "bonds2": fx["bonds"][1], "bonds3": fx["bonds"][2],
            "name": "GEOM" if fx["limit_bonds_to_one"] else "QM9"}
    bonds, e, poff = bond_orders_batch(fx["x"].cuda(), fx["atom_types"].cuda(), torch.tensor(fx["sizes"]), info, fx["margins"])
    rows = []
    for k, e_ref in enumerate(fx["bond_E"]):
        n = fx["sizes"][k]
        got = e[int(poff[k]): int(poff[k]) + n * n].reshape(n, n).cpu()
        assert torch.equal(got, e_ref), k
        nz = torch.nonzero(e_ref)                       # the reference's `all_bonds` loop order
        rows += [(k, int(i), int(j), int(e_ref[i, j])) for i, j in nz.tolist()]
    assert bonds.cpu().tolist() == [list(r) for r in rows]
