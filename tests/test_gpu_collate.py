"""GPU: bdiff.collate.PackedDataset.collate (two kernels) against the reference batch restricted to mask == True rows
(tests/golden/collate.pt) — bit-exact — and the denoiser accepts the packed batch."""
import os

import pytest
import torch

from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_packed_collation_matches_reference_bit_exact():
    from bdiff.collate import PackedDataset
    fx = torch.load(os.path.join(GOLDEN, "collate.pt"), weights_only=False)
    ds = PackedDataset({k: fx[k] for k in ("positions", "charges", "one_hot", "alpha", "mu")}, torch.device("cuda"),
                       properties=("alpha", "mu"))
    b = ds.collate(fx["idx"], conditioning=("alpha", "mu"), property_norms=fx["norms"])
    ref = fx["ref"]
    assert torch.equal(b.x.cpu(), ref["x"]) and torch.equal(b.one_hot.cpu(), ref["one_hot"])
    assert torch.equal(b.charges.cpu(), ref["charges"]) and torch.equal(b.batch.cpu(), ref["batch"])
    assert torch.equal(b.props_context.cpu(), ref["context"])
    assert b.mask.all() and b.num_nodes == ref["present_rows"] and b.num_graphs == len(fx["idx"])
    assert b.num_nodes_present.cpu().tolist() == torch.bincount(ref["batch"]).tolist()
    b2 = ds.collate(fx["idx"][:3], conditioning=("mu",), property_norms=fx["norms"])
    assert torch.equal(b2.props_context.cpu()[:, 0], ref["context"][: b2.num_nodes, 1])


def test_denoiser_runs_on_a_packed_batch():
    import bdiff
    import gcpnet_oracle as O
    from bdiff.collate import PackedDataset
    fx = torch.load(os.path.join(GOLDEN, "collate.pt"), weights_only=False)
    ds = PackedDataset({k: fx[k] for k in ("positions", "charges", "one_hot", "alpha")}, torch.device("cuda"), properties=("alpha",))
    b = ds.collate(fx["idx"], conditioning=("alpha",), property_norms=fx["norms"])
    ocfg = O.config_named("qm9_cond")
    sd = O.random_state_dict(ocfg, 3)
    net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named("qm9_cond"), mode="tensor")
    net.load_state_dict(sd, strict=True)
    net.cuda()
    _, xc = O.centralize(b.x.cpu(), b.batch.cpu(), b.mask.cpu(), b.num_graphs)
    xh = torch.cat((xc, b.one_hot.cpu()), -1).cuda()             # qm9_cond: 5 atom types, no charges
    t = torch.full((b.num_nodes, 1), 0.5, device="cuda")
    with torch.no_grad():
        _, out = net(b, xh, t)
    ref = O.denoiser_forward(sd, ocfg, b.batch.cpu(), b.mask.cpu(), xh.cpu(), t.cpu(), b.props_context.cpu())
    assert (out.cpu() - ref).abs().max().item() <= 1e-4 * max(1.0, ref.abs().max().item())
