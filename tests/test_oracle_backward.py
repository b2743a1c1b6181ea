"""CPU: the hand-derived backward of the denoiser (oracle/gcpnet_backward.py — the formulas a CUDA backward has to
implement) against torch.autograd through the forward oracle, for every parameter tensor; masked atoms, a one-atom
molecule and both shipped configurations included."""
import pytest
import torch

import gcpnet_backward as B
import gcpnet_oracle as O


@pytest.mark.parametrize("cname,sizes", [("qm9", [5, 1, 7]), ("qm9_cond", [4, 6]), ("geom", [9, 3])])
def test_manual_backward_matches_autograd(cname, sizes):
    cfg = O.config_named(cname)
    sd = O.random_state_dict(cfg, 21, scale=0.7)
    g = torch.Generator().manual_seed(5)
    nmol = len(sizes)
    bi = torch.repeat_interleave(torch.arange(nmol), torch.tensor(sizes))
    n = bi.shape[0]
    mask = torch.ones(n, dtype=torch.bool)
    mask[2] = False
    xh = torch.randn((n, 3 + cfg.num_h), generator=g)
    t = torch.full((n, 1), 0.37)
    ctx = torch.randn((n, cfg.num_context), generator=g) if cfg.num_context else None
    d_out = torch.randn((n, 3 + cfg.num_h), generator=g)
    # reference gradients: autograd through the forward oracle
    sda = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
    out_a = O.denoiser_forward(sda, cfg, bi, mask, xh, t, ctx)
    (out_a * d_out).sum().backward()
    # manual
    out_m, tape = B.denoiser_forward_with_tape(sd, cfg, bi, mask, xh, t, ctx)
    assert torch.allclose(out_m, out_a.detach(), rtol=1e-5, atol=1e-6)
    grads = B.denoiser_backward(sd, cfg, bi, mask, tape, d_out)
    assert set(grads.keys()) == set(sd.keys())
    worst, worst_key = 0.0, None
    for k in sd:
        ref = sda[k].grad
        assert ref is not None, k
        err = (grads[k] - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)
        if err > worst:
            worst, worst_key = err, k
    assert worst < 2e-4, (worst_key, worst)
