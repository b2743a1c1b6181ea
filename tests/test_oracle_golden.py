"""CPU: the oracle restatement reproduces the reference outputs stored in tests/golden/.

The fixtures were produced by the unmodified reference (tests/golden/make_golden.py); the weights are
regenerated from their seed and checksummed first.  Tolerances: integer work bit-exact; floating
point max-abs <= 2e-5 * max(1, |ref|_max) per forward (the measured fp32-vs-fp64 floor is ~1e-6,
SURVEY.md §8c), chains 1e-4 relative (chain fixtures use 0.5-scaled weights: well-conditioned regime).
"""
import pytest
import torch

import gcpnet_oracle as O
from conftest import load_golden

FWD_CASES = ["qm9_small_masked", "qm9_tiny_sizes", "qm9_b4_n19", "qm9_cond", "geom_mixed", "geom_max181"]


def weights_for(fx):
    cfg = O.config_named(fx["config"])
    sd = O.random_state_dict(cfg, fx["weight_seed"], scale=fx.get("weight_scale", 1.0))
    tot = sum(float(v.double().sum()) for v in sd.values())
    sq = sum(float((v.double() ** 2).sum()) for v in sd.values())
    assert abs(tot - float(fx["weight_checksum"][0])) < 1e-6 * max(1.0, abs(tot))
    assert abs(sq - float(fx["weight_checksum"][1])) < 1e-9 * sq
    return cfg, sd


def test_edge_index_kat():
    fx = load_golden("kat_edge_index")
    ei = O.fully_connected_edge_index(fx["batch_index"], fx["mask"])
    assert ei.dtype == torch.int64 and torch.equal(ei, fx["edge_index"])
    assert ei[0].tolist() == [0, 0, 0, 1, 1, 1, 2, 2, 2] + [5] * 4 + [6] * 4 + [7] * 4 + [8] * 4
    assert torch.equal(O.fully_connected_edge_index(fx["batch_index"], None), fx["edge_index_nomask"])


def test_edge_index_empty_and_single():
    bi = torch.zeros(3, dtype=torch.int64)
    assert O.fully_connected_edge_index(bi, torch.zeros(3, dtype=torch.bool)).shape == (2, 0)
    ei = O.fully_connected_edge_index(torch.zeros(1, dtype=torch.int64), None)
    assert ei.tolist() == [[0], [0]]


@pytest.mark.parametrize("name", FWD_CASES)
def test_forward_matches_reference(name):
    fx = load_golden(name)
    cfg, sd = weights_for(fx)
    taps = {}
    out = O.denoiser_forward(sd, cfg, fx["batch_index"], fx["mask"], fx["xh"], fx["t"], fx["context"], taps=taps)
    ref = fx["net_out"]
    tol = 2e-5 * max(1.0, ref.abs().max().item())
    assert (out - ref).abs().max().item() <= tol
    assert taps["edge_index"].shape[1] == fx["num_edges"]
    if "edge_index" in fx:
        assert torch.equal(taps["edge_index"], fx["edge_index"])
        assert (taps["f_ij"] - fx["f_ij"]).abs().max().item() <= 1e-6
        assert (taps["e"] - fx["e"]).abs().max().item() <= 1e-5
        assert (taps["xi"] - fx["xi"]).abs().max().item() <= 1e-5
        for mine, theirs in zip(taps["layers"], fx["layers"]):
            for k in ("h", "chi", "x"):
                scale = max(1.0, theirs[k].abs().max().item())
                assert (mine[k] - theirs[k]).abs().max().item() <= 2e-5 * scale, k


@pytest.mark.parametrize("name", ["chain_qm9_T6", "chain_qm9_cond_T4", "chain_geom_T3"])
def test_chain_matches_reference(name):
    fx = load_golden(name)
    cfg, sd = weights_for(fx)
    assert torch.equal(O.gamma_table(cfg.num_timesteps, cfg.noise_precision, cfg.schedule_power), fx["gamma"])
    torch.manual_seed(fx["noise_seed"])
    out, bi, mask, z0 = O.sample_chain(sd, cfg, torch.tensor(fx["sizes"]), lambda s: torch.randn(s),
                                       num_timesteps=fx["steps"], context=fx["context"], return_z0=True)
    rel = (z0 - fx["z_0"]).abs().max().item() / fx["z_0"].abs().max().item()
    assert rel < 1e-4
    a = cfg.num_atom_types
    assert torch.equal(out[:, 3:3 + a], fx["out"][:, 3:3 + a])
    relx = (out[:, :3] - fx["out"][:, :3]).abs().max().item() / fx["out"][:, :3].abs().max().item()
    assert relx < 1e-4


def test_se3_equivariance_of_oracle():
    """Rotation + translation of x rotates vel and leaves h invariant (reflection is NOT a symmetry)."""
    cfg = O.config_named("qm9")
    sd = O.random_state_dict(cfg, 3)
    g = torch.Generator().manual_seed(5)
    sizes = [6, 4]
    bi = torch.repeat_interleave(torch.arange(2), torch.tensor(sizes))
    mask = torch.ones(10, dtype=torch.bool)
    xh = torch.randn((10, 9), generator=g)
    _, xc = O.centralize(xh[:, :3], bi, mask, 2)
    t = torch.full((10, 1), 0.3)
    A = torch.randn((3, 3), generator=g)
    Q, _ = torch.linalg.qr(A)
    if torch.det(Q) < 0:
        Q[:, 0] = -Q[:, 0]
    o1 = O.denoiser_forward(sd, cfg, bi, mask, torch.cat((xc, xh[:, 3:]), -1), t, dtype=torch.float64)
    o2 = O.denoiser_forward(sd, cfg, bi, mask, torch.cat((xc @ Q.T, xh[:, 3:]), -1), t, dtype=torch.float64)
    assert (o2[:, :3] - o1[:, :3] @ Q.double().T).abs().max() < 1e-6
    assert (o2[:, 3:] - o1[:, 3:]).abs().max() < 1e-6


@pytest.mark.parametrize("name", ["nll_qm9", "nll_geom"])
def test_eval_nll_matches_reference(name):
    """Evaluation-mode NLL terms (two denoiser calls) vs the reference; same global RNG stream."""
    fx = load_golden(name)
    cfg, sd = weights_for(fx)
    torch.manual_seed(fx["rng_seed"])
    nll, terms = O.eval_nll(sd, cfg, fx["batch_index"], fx["mask"], fx["x"], fx["one_hot"], fx["charges"], None,
                            fx["histogram"], lambda s: torch.randn(s))
    assert torch.equal(terms["t_int"], fx["terms"]["t_int"])
    for k, ref in fx["terms"].items():
        if k == "t_int":
            continue
        assert torch.allclose(terms[k], ref, rtol=1e-5, atol=1e-5), k
    assert torch.allclose(nll, fx["nll"], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("name", ["train_qm9", "train_geom"])
def test_training_loss_matches_reference(name):
    """Training-mode L2 objective (one denoiser call, t == 0 molecule included) vs the reference in .train() mode; the
    fixture's t_int was injected into the reference, the noise comes from the same global RNG stream."""
    fx = load_golden(name)
    cfg, sd = weights_for(fx)
    torch.manual_seed(fx["rng_seed"])
    loss, terms = O.eval_nll(sd, cfg, fx["batch_index"], fx["mask"], fx["x"], fx["one_hot"], fx["charges"], None,
                             fx["histogram"], lambda s: torch.randn(s), t_int=fx["terms"]["t_int"].reshape(-1, 1),
                             training=True)
    assert (fx["terms"]["t_int"] == 0).any()
    for k, ref in fx["terms"].items():
        if k == "t_int":
            continue
        assert torch.allclose(terms[k], ref, rtol=1e-5, atol=1e-5), k
    assert torch.allclose(loss, fx["nll"], rtol=1e-5, atol=1e-4)


@pytest.mark.parametrize("name", ["optimize_qm9_cond_T4", "optimize_geom_T3"])
def test_optimize_chain_matches_reference(name):
    """mol_gen_optimize of the reference (chain started from given samples) vs the oracle, same global RNG stream."""
    fx = load_golden(name)
    cfg, sd = weights_for(fx)
    nmol = len(fx["sizes"])
    num_nodes = torch.tensor(fx["sizes"])
    bi = torch.repeat_interleave(torch.arange(nmol), num_nodes)
    mask = torch.ones(bi.shape[0], dtype=torch.bool)
    z = O.normalize_samples(cfg, fx["x"], fx["one_hot"], mask)
    torch.manual_seed(fx["noise_seed"])
    out, _, _ = O.sample_chain(sd, cfg, num_nodes, lambda s: torch.randn(s), num_timesteps=fx["steps"], context=fx["context"],
                               z_init=z)
    a = cfg.num_atom_types
    assert torch.equal(out[:, 3:3 + a], fx["out"][:, 3:3 + a])
    rel = (out[:, :3] - fx["out"][:, :3]).abs().max().item() / fx["out"][:, :3].abs().max().item()
    assert rel < 1e-5, rel


@pytest.mark.parametrize("name", ["grad_geom", "grad_qm9"])
def test_training_gradients_match_reference(name):
    """Config 5's parity target for the (not yet built) CUDA backward: autograd through the ORACLE's training objective
    gives the same gradient for every denoiser parameter as loss.backward() through the unmodified reference
    (fingerprints: L2 norm, sum and 8 strided entries per tensor; 202 / 432 tensors, none without gradient)."""
    fx = load_golden(name)
    cfg, sd = weights_for(fx)
    sd = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    torch.manual_seed(fx["rng_seed"])
    loss, terms = O.eval_nll(sd, cfg, fx["batch_index"], fx["mask"], fx["x"], fx["one_hot"], fx["charges"], None,
                             fx["histogram"], lambda s: torch.randn(s), t_int=fx["terms"]["t_int"].reshape(-1, 1),
                             training=True)
    assert torch.allclose(loss.detach(), fx["nll"], rtol=1e-5, atol=1e-4)
    loss.mean().backward()
    assert set(fx["grads"].keys()) == set(sd.keys())
    worst = 0.0
    for k, ref in fx["grads"].items():
        g = sd[k].grad
        assert g is not None, k
        f = g.detach().double().reshape(-1)
        scale = max(ref["norm"], 1e-12)
        worst = max(worst, abs(float(f.norm()) - ref["norm"]) / scale, abs(float(f.sum()) - ref["sum"]) / scale,
                    float((f[ref["idx"]].float() - ref["vals"]).abs().max()) / scale)
    assert worst < 2e-4, worst
