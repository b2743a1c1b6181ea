"""Generate the golden fixtures in tests/golden/*.pt from the UNMODIFIED reference.

Build-container only: imports /root/reference through oracle/ref_shim.py.  The fixtures hold inputs
and the reference's outputs (and intermediate taps); the WEIGHTS are not stored — they are regenerated
anywhere from `oracle.gcpnet_oracle.random_state_dict(cfg, seed)` (torch CPU generator, deterministic)
and were loaded into the reference module with `load_state_dict(strict=True)` before it was run.  A
checksum of the weights is stored so a drift of the generator is caught before any comparison.

Run:  python tests/golden/make_golden.py
"""
import os
import sys
import warnings

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
warnings.filterwarnings("ignore")
import ref_shim  # noqa: E402
import gcpnet_oracle as O  # noqa: E402

WEIGHT_SEED = 7


def weight_checksum(sd):
    tot = sum(float(v.double().sum()) for v in sd.values())
    sq = sum(float((v.double() ** 2).sum()) for v in sd.values())
    return torch.tensor([tot, sq], dtype=torch.float64)


def make_inputs(cfg, sizes, seed, mask_mode="none"):
    g = torch.Generator().manual_seed(seed)
    B = len(sizes)
    bi = torch.repeat_interleave(torch.arange(B), torch.tensor(sizes))
    N = bi.shape[0]
    mask = torch.ones(N, dtype=torch.bool)
    if mask_mode == "pad":   # suffix padding on even molecules + one interior hole
        off = 0
        for k, n in enumerate(sizes):
            if k % 2 == 0 and n > 3:
                mask[off + n - 2: off + n] = False
            off += n
        mask[1] = False
    xh = torch.randn((N, 3 + cfg.num_h), generator=g) * mask[:, None]
    _, xc = O.centralize(xh[:, :3], bi, mask, B)
    xh = torch.cat((xc, xh[:, 3:]), -1)
    t = torch.rand((B, 1), generator=g)[bi].contiguous()
    ctx = (torch.randn((B, cfg.num_context), generator=g)[bi] * mask[:, None]) if cfg.num_context else None
    return bi, mask, xh, t, ctx


def reference_with_weights(cname):
    net, _ = ref_shim.build_reference_dynamics(cname, seed=0)
    cfg = O.config_named(cname)
    sd = O.random_state_dict(cfg, WEIGHT_SEED)
    net.load_state_dict(sd, strict=True)
    return net, cfg, sd


def forward_case(cname, sizes, seed, mask_mode="none", layer_taps=False):
    net, cfg, sd = reference_with_weights(cname)
    bi, mask, xh, t, ctx = make_inputs(cfg, sizes, seed, mask_mode)
    batch = ref_shim.Batch(batch=bi, mask=mask, props_context=ctx)
    taps = []
    hooks = []
    if layer_taps:
        for layer in net.interaction_layers:
            hooks.append(layer.register_forward_hook(
                lambda m, i, o: taps.append(dict(h=o[0][0].clone(), chi=o[0][1].clone(), x=o[1].clone()))))
    with torch.no_grad():
        _, out = net(batch, xh.clone(), t)
    for h in hooks:
        h.remove()
    fx = dict(config=cname, sizes=list(sizes), weight_seed=WEIGHT_SEED, weight_checksum=weight_checksum(sd),
              batch_index=bi, mask=mask, xh=xh, t=t, context=ctx, net_out=out.clone(),
              num_edges=int(batch.edge_index.shape[1]))
    if layer_taps:
        fx.update(edge_index=batch.edge_index.clone(), f_ij=batch.f_ij.clone(), e=batch.e.clone(),
                  xi=batch.xi.clone(), x_final=batch.x.clone(), chi_final=batch.chi.clone(), layers=taps)
    return fx


def chain_case(cname, sizes, steps, seed, scale=0.5):
    """Short sampling chain.  Weights are scaled by 0.5: with full-size random weights the untrained denoiser
    drives |z| to 1e5..1e7 within a few strided steps and the chain amplifies fp32 round-off to ~1e-4..1e-2
    (chaotic regime); at 0.5 the chain is well-conditioned (fp32-vs-fp64 ~3e-7) and can be pinned tightly."""
    ddpm, _ = ref_shim.build_reference_ddpm(cname, seed=0)
    cfg = O.config_named(cname)
    sd = O.random_state_dict(cfg, WEIGHT_SEED, scale=scale)
    ddpm.dynamics_network.load_state_dict(sd, strict=True)
    zs = []
    hook = ddpm.dynamics_network.register_forward_pre_hook(lambda m, args: zs.append(args[1].clone()))
    num_nodes = torch.tensor(sizes)
    ctx = None
    if cfg.num_context:
        ctx = torch.randn((len(sizes), cfg.num_context), generator=torch.Generator().manual_seed(seed + 1))
    torch.manual_seed(seed)
    out, bi, mask = ddpm.mol_gen_sample(num_samples=len(sizes), num_nodes=num_nodes, device="cpu",
                                        num_timesteps=steps, context=ctx)
    hook.remove()
    return dict(config=cname, sizes=list(sizes), steps=steps, noise_seed=seed, weight_seed=WEIGHT_SEED,
                weight_scale=scale, weight_checksum=weight_checksum(sd), context=ctx, out=out.clone(), z_T=zs[0], z_1=zs[1],
                z_0=zs[-1], gamma=ddpm.gamma.gamma.data.clone())


def optimize_case(cname, sizes, steps, seed, scale=0.5):
    """mol_gen_optimize (variational_diffusion.py:1414-1546) of the reference on given samples (unmodified)."""
    ddpm, _ = ref_shim.build_reference_ddpm(cname, seed=0)
    cfg = O.config_named(cname)
    assert not cfg.include_charges
    sd = O.random_state_dict(cfg, WEIGHT_SEED, scale=scale)
    ddpm.dynamics_network.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(seed)
    nmol = len(sizes)
    num_nodes = torch.tensor(sizes)
    bi = torch.repeat_interleave(torch.arange(nmol), num_nodes)
    n = bi.shape[0]
    mask = torch.ones(n, dtype=torch.bool)
    x = torch.randn((n, 3), generator=g) * 1.5
    _, x = O.centralize(x, bi, mask, nmol)
    one_hot = torch.nn.functional.one_hot(torch.randint(0, cfg.num_atom_types, (n,), generator=g), cfg.num_atom_types).float()
    ctx = torch.randn((nmol, cfg.num_context), generator=g) if cfg.num_context else None
    samples, o = [], 0
    for k in sizes:
        samples.append((x[o:o + k].clone(), one_hot[o:o + k].clone()))
        o += k
    torch.manual_seed(seed)
    out, bi2, mask2 = ddpm.mol_gen_optimize(samples=samples, num_nodes=num_nodes, device="cpu", num_timesteps=steps,
                                            context=ctx)
    assert torch.equal(bi2, bi)
    return dict(config=cname, sizes=list(sizes), steps=steps, noise_seed=seed, weight_seed=WEIGHT_SEED, weight_scale=scale,
                weight_checksum=weight_checksum(sd), context=ctx, x=x, one_hot=one_hot, out=out.clone())


def grad_case(cname, sizes, seed, t_fixed, scale=0.5):
    """Config 5's parity target: mean training loss of the reference in .train() mode and its gradient with respect to
    every denoiser parameter (loss.backward() through the unmodified reference).  The gradients are stored as
    per-tensor fingerprints (L2 norm, sum, and the 8 entries at fixed strided positions) to keep the fixture small."""
    fx = train_case(cname, sizes, seed, t_fixed, scale, _with_grad=True)
    return fx


def nll_case(cname, sizes, seed, scale=0.5):
    """Evaluation-mode NLL terms of the reference (EquivariantVariationalDiffusion.forward with .eval(), two denoiser
    calls) + the Lightning module's evaluation assembly restated from qm9_mol_gen_ddpm.py:247-262."""
    hist = {19: 5, 7: 2, 12: 3, 30: 1, 6: 1, 18: 2, 44: 3, 43: 2, 25: 1}
    ddpm, _ = ref_shim.build_reference_ddpm(cname, seed=0, n_nodes_hist=hist)
    cfg = O.config_named(cname)
    sd = O.random_state_dict(cfg, WEIGHT_SEED, scale=scale)
    ddpm.dynamics_network.load_state_dict(sd, strict=True)
    ddpm.eval()
    g = torch.Generator().manual_seed(seed)
    nmol = len(sizes)
    bi = torch.repeat_interleave(torch.arange(nmol), torch.tensor(sizes))
    n = bi.shape[0]
    mask = torch.ones(n, dtype=torch.bool)
    mask[sizes[0] + 1] = False
    x = torch.randn((n, 3), generator=g) * mask[:, None]
    _, x = O.centralize(x, bi, mask, nmol)
    types = torch.randint(0, cfg.num_atom_types, (n,), generator=g)
    one_hot = torch.nn.functional.one_hot(types, cfg.num_atom_types).float() * mask[:, None]
    charges = (torch.randint(1, 9, (n,), generator=g).float() * mask) if cfg.include_charges else torch.zeros((n, 0))
    num_present = torch.zeros(nmol, dtype=torch.long).index_add_(0, bi, mask.long())
    batch = ref_shim.Batch(batch=bi, mask=mask, x=x.clone(), h={"categorical": one_hot.clone(), "integer": charges.clone()},
                           num_graphs=nmol, num_nodes_present=num_present, props_context=None)
    torch.manual_seed(seed + 1000)
    with torch.no_grad():
        (dlp, err, snr, l0x, l0h, nlc, klp, lpn, tint, _info) = ddpm(batch, return_loss_info=True)
    T = cfg.num_timesteps
    nll = T * 0.5 * snr * err + (l0x + l0h + nlc) + klp - dlp - lpn
    return dict(config=cname, sizes=list(sizes), weight_seed=WEIGHT_SEED, weight_scale=scale,
                weight_checksum=weight_checksum(sd), histogram=hist, rng_seed=seed + 1000, batch_index=bi, mask=mask, x=x,
                one_hot=one_hot, charges=charges, nll=nll.clone(),
                terms=dict(delta_log_px=dlp, error_t=err, SNR_weight=snr, loss_0_x=l0x, loss_0_h=l0h,
                           neg_log_constants=nlc, kl_prior=klp, log_pN=lpn, t_int=tint))


def grad_fingerprint(g):
    f = g.detach().double().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, steps=min(8, f.numel())).long()
    return dict(norm=float(f.norm()), sum=float(f.sum()), idx=idx, vals=f[idx].float().clone())


def train_case(cname, sizes, seed, t_fixed, scale=0.5, _with_grad=False):
    """Training-mode L2 objective of the reference (EquivariantVariationalDiffusion.forward with .train(): one denoiser
    call) + the Lightning module's training assembly restated from qm9_mol_gen_ddpm.py:232-262.  torch.randint is
    patched for the duration of the call so that the fixture contains a t == 0 molecule (the L0 branch)."""
    hist = {19: 5, 7: 2, 12: 3, 30: 1, 6: 1, 18: 2, 44: 3, 43: 2, 25: 1}
    ddpm, _ = ref_shim.build_reference_ddpm(cname, seed=0, n_nodes_hist=hist)
    cfg = O.config_named(cname)
    sd = O.random_state_dict(cfg, WEIGHT_SEED, scale=scale)
    ddpm.dynamics_network.load_state_dict(sd, strict=True)
    ddpm.train()
    g = torch.Generator().manual_seed(seed)
    nmol = len(sizes)
    bi = torch.repeat_interleave(torch.arange(nmol), torch.tensor(sizes))
    n = bi.shape[0]
    mask = torch.ones(n, dtype=torch.bool)
    mask[sizes[0] + 1] = False
    x = torch.randn((n, 3), generator=g) * mask[:, None]
    _, x = O.centralize(x, bi, mask, nmol)
    types = torch.randint(0, cfg.num_atom_types, (n,), generator=g)
    one_hot = torch.nn.functional.one_hot(types, cfg.num_atom_types).float() * mask[:, None]
    charges = (torch.randint(1, 9, (n,), generator=g).float() * mask) if cfg.include_charges else torch.zeros((n, 0))
    num_present = torch.zeros(nmol, dtype=torch.long).index_add_(0, bi, mask.long())
    batch = ref_shim.Batch(batch=bi, mask=mask, x=x.clone(), h={"categorical": one_hot.clone(), "integer": charges.clone()},
                           num_graphs=nmol, num_nodes_present=num_present, props_context=None)
    torch.manual_seed(seed + 2000)
    t_fixed = torch.tensor(t_fixed, dtype=torch.long).reshape(nmol, 1)
    real_randint = torch.randint
    torch.randint = lambda *a, **k: t_fixed.clone()
    try:
        with torch.set_grad_enabled(_with_grad):
            (dlp, err, snr, l0x, l0h, nlc, klp, lpn, tint, _info) = ddpm(batch, return_loss_info=True)
    finally:
        torch.randint = real_randint
    denom = (3 + cfg.num_h) * num_present.float()                      # norm_training_by_max_nodes: false
    loss = 0.5 * (err / denom) + (l0x / denom + l0h) + klp - dlp - lpn
    grads = None
    if _with_grad:
        loss.mean().backward()                                         # training_step: loss = nll.mean(0)
        grads = {k: grad_fingerprint(p.grad) for k, p in ddpm.dynamics_network.named_parameters()}
        assert all(p.grad is not None for p in ddpm.dynamics_network.parameters())
        loss, dlp, err, snr, l0x, l0h, nlc, klp, lpn = (v.detach() for v in (loss, dlp, err, snr, l0x, l0h, nlc, klp, lpn))
    return dict(config=cname, sizes=list(sizes), weight_seed=WEIGHT_SEED, weight_scale=scale,
                weight_checksum=weight_checksum(sd), histogram=hist, rng_seed=seed + 2000, batch_index=bi, mask=mask, x=x,
                one_hot=one_hot, charges=charges, nll=loss.clone(), grads=grads,
                terms=dict(delta_log_px=dlp, error_t=err, SNR_weight=snr, loss_0_x=l0x, loss_0_h=l0h,
                           neg_log_constants=nlc, kl_prior=klp, log_pN=lpn, t_int=tint))


def main():
    only = sys.argv[1:]        # optional: names of the fixtures to (re)generate
    ref_shim.install()
    fixtures = {
        # integer KAT for the edge index, observed from the reference
        "kat_edge_index": None,       # filled below
        "qm9_small_masked": lambda: forward_case("qm9", [5, 9, 3, 7], 11, "pad", layer_taps=True),
        "qm9_tiny_sizes": lambda: forward_case("qm9", [1, 2, 3, 1], 12),
        "qm9_b4_n19": lambda: forward_case("qm9", [19, 19, 19, 19], 13),
        "qm9_cond": lambda: forward_case("qm9_cond", [19, 12, 23], 14),
        "geom_mixed": lambda: forward_case("geom", [44, 30, 61, 25], 15),
        "geom_max181": lambda: forward_case("geom", [181, 3], 16),
        "chain_qm9_T6": lambda: chain_case("qm9", [19, 7, 12], 6, 123),
        "chain_qm9_cond_T4": lambda: chain_case("qm9_cond", [9, 14], 4, 321),
        "chain_geom_T3": lambda: chain_case("geom", [30, 44], 3, 77),
        "nll_qm9": lambda: nll_case("qm9", [19, 7, 12], 5),
        "nll_geom": lambda: nll_case("geom", [30, 44, 25], 6),
        "train_qm9": lambda: train_case("qm9", [19, 7, 12], 8, [0, 517, 1000]),
        "train_geom": lambda: train_case("geom", [30, 44, 25], 9, [311, 0, 42]),
        "grad_geom": lambda: grad_case("geom", [12, 8, 6], 10, [311, 0, 42]),
        "grad_qm9": lambda: grad_case("qm9", [7, 7], 11, [100, 900]),
        "optimize_qm9_cond_T4": lambda: optimize_case("qm9_cond", [9, 14, 19], 4, 41),
        "optimize_geom_T3": lambda: optimize_case("geom", [30, 21], 3, 42),
    }
    from src.models.components.gcpnet import GCPNetDynamics
    bi = torch.tensor([0] * 5 + [1] * 5)
    mk = torch.tensor([1, 1, 1, 0, 0, 1, 1, 1, 1, 0], dtype=torch.bool)
    fixtures["kat_edge_index"] = lambda: dict(
        batch_index=bi, mask=mk,
        edge_index=GCPNetDynamics.get_fully_connected_edge_index(bi, mk),
        edge_index_nomask=GCPNetDynamics.get_fully_connected_edge_index(bi, None))
    for name, make in fixtures.items():
        if only and name not in only:
            continue
        fx = make()
        path = os.path.join(HERE, name + ".pt")
        torch.save(fx, path)
        print(f"{name:22s} {os.path.getsize(path) / 1024:8.1f} KiB")


if __name__ == "__main__":
    main()
