"""Validate the CPU restatement (gcpnet_oracle.py) against the UNMODIFIED reference modules.

Build-container only (needs /root/reference via ref_shim).  Run: python oracle/check_against_reference.py
Compares, for each shipped config: parameter names/shapes, edge_index (bit-exact), every geometry
tensor, embedding outputs, per-layer (h, chi, x), net_out, the gamma table, and a short sampling chain
with shared noise.  Prints max-abs differences and exits non-zero on a violation.
"""
import os
import sys

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import gcpnet_oracle as O  # noqa: E402


def make_inputs(cfg, sizes, seed=123, masked=False):
    g = torch.Generator().manual_seed(seed)
    num_nodes = torch.tensor(sizes)
    B = len(sizes)
    bi = torch.repeat_interleave(torch.arange(B), num_nodes)
    N = bi.shape[0]
    mask = torch.ones(N, dtype=torch.bool)
    if masked:  # QM9-training-like suffix padding + one interior hole
        off = 0
        for k, n in enumerate(sizes):
            if k % 2 == 0 and n > 3:
                mask[off + n - 2: off + n] = False
            off += n
        mask[1] = False
    xh = torch.randn((N, 3 + cfg.num_h), generator=g) * mask[:, None]
    _, xc = O.centralize(xh[:, :3], bi, mask, B)
    xh = torch.cat((xc, xh[:, 3:]), -1)
    t = torch.rand((B, 1), generator=g)[bi]
    ctx = torch.randn((B, cfg.num_context), generator=g)[bi] * mask[:, None] if cfg.num_context else None
    return bi, mask, xh, t, ctx


def check(name, a, b, tol):
    d = (a.double() - b.double()).abs().max().item() if a.numel() else 0.0
    ok = d <= tol
    print(f"  {name:28s} max|diff| = {d:.3e}  (tol {tol:.1e})  {'ok' if ok else 'FAIL'}")
    return ok


def main():
    ok = True
    for cname, sizes, masked in (("qm9", [19, 19, 19, 19], False), ("qm9", [5, 9, 3, 12, 7], True),
                                 ("qm9_cond", [19, 12, 23], False), ("geom", [44, 30, 61, 25], False)):
        print(f"== {cname} sizes={sizes} masked={masked}")
        net, _ = ref_shim.build_reference_dynamics(cname, seed=0)
        cfg = O.config_named(cname)
        sd = {k: v.detach().clone() for k, v in net.state_dict().items()}
        shapes = O.param_shapes(cfg)
        assert set(shapes) == set(sd), set(shapes) ^ set(sd)
        assert all(tuple(sd[k].shape) == shapes[k] for k in sd)
        bi, mask, xh, t, ctx = make_inputs(cfg, sizes, masked=masked)
        batch = ref_shim.Batch(batch=bi, mask=mask, props_context=ctx)
        with torch.no_grad():
            _, ref_out = net(batch, xh, t)
        taps = {}
        out = O.denoiser_forward(sd, cfg, bi, mask, xh, t, ctx, taps=taps)
        ok &= bool(torch.equal(taps["edge_index"], batch.edge_index)); print("  edge_index bit-exact:", torch.equal(taps["edge_index"], batch.edge_index), tuple(batch.edge_index.shape))
        ok &= check("f_ij", taps["f_ij"], batch.f_ij, 1e-6)
        ok &= check("e (embedded)", taps["e"], batch.e, 2e-6)
        ok &= check("xi (embedded)", taps["xi"], batch.xi, 2e-6)
        ok &= check("chi (final)", taps["layers"][-1]["chi"], batch.chi, 5e-5)
        ok &= check("net_out", out, ref_out, 2e-5)
        out64 = O.denoiser_forward(sd, cfg, bi, mask, xh, t, ctx, dtype=torch.float64)
        check("net_out fp32 vs oracle fp64", out, out64.float(), 1e-4)
        check("reference vs oracle fp64", ref_out, out64.float(), 1e-4)

    print("== gamma table")
    ddpm, _ = ref_shim.build_reference_ddpm("qm9", seed=0)
    ok &= bool(torch.equal(ddpm.gamma.gamma.data, O.gamma_table())); print("  bit-exact:", torch.equal(ddpm.gamma.gamma.data, O.gamma_table()))

    print("== sampling chain qm9 B=3 T=8 (shared RNG stream)")
    cfg = O.config_named("qm9")
    sd = {k: v.detach().clone() for k, v in ddpm.dynamics_network.state_dict().items()}
    num_nodes = torch.tensor([19, 7, 12])
    torch.manual_seed(123)
    ref, rbi, rmask = ddpm.mol_gen_sample(num_samples=3, num_nodes=num_nodes, device="cpu", num_timesteps=8)
    torch.manual_seed(123)
    mine, bi, mask = O.sample_chain(sd, cfg, num_nodes, lambda shape: torch.randn(shape), num_timesteps=8)
    rel = (mine[:, :3] - ref[:, :3]).abs().max().item() / ref[:, :3].abs().max().item()
    types_eq = torch.equal(mine[:, 3:8], ref[:, 3:8])
    # the untrained net blows the charge channel up to ~1e5; compare it relatively
    rel_q = (mine[:, 8] - ref[:, 8]).abs().max().item() / ref[:, 8].abs().max().clamp(min=1).item()
    print(f"  x rel diff {rel:.3e}; atom types equal: {types_eq}; charge rel diff {rel_q:.3e}")
    ok &= rel < 1e-3 and types_eq and rel_q < 1e-3
    print("ALL OK" if ok else "FAILURES")
    sys.exit(0 if ok else 1)


if __name__ == "__main__":
    main()
