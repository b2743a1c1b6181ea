"""GPU tests of the tensor-core (tcgen05) path.

Tolerances (tensor mode = bf16 operands, fp32 accumulation in TMEM, tanh.approx activations): the hardware
self test is compared with a bf16-rounded fp32 matmul at 1e-3 relative; a full denoiser forward is compared with
the reference golden output at max-abs <= 2e-2 * max(1,|ref|) and rms <= 5e-3 * rms(ref)... (SURVEY.md §8c:
emulated bf16/TF32 operand rounding of the reference itself gives max-abs 6e-3, rms 1.6e-3).
"""
import ctypes as C

import pytest
import torch

import gcpnet_oracle as O
from conftest import load_golden

pytestmark = pytest.mark.gpu


def test_umma_selftest_matches_bf16_matmul():
    import bdiff
    lib = bdiff.load_library()
    g = torch.Generator().manual_seed(0)
    a = torch.randn((128, 128), generator=g).cuda()
    w = torch.randn((320, 128), generator=g).cuda()
    c = torch.zeros((128, 328), device="cuda")
    rc = lib.bdiff_selftest_umma(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(a.data_ptr()),
                                 C.c_void_p(w.data_ptr()), C.c_void_p(c.data_ptr()))
    assert rc == 0
    ref = a.bfloat16().float() @ w.bfloat16().float().t()
    ref[:, 288:] = -ref[:, 288:]
    err = (c[:, :320] - ref).abs().max().item() / ref.abs().max().item()
    assert err < 1e-3, f"UMMA self test rel err {err:.3e}"
    scratch = torch.arange(128 * 8, device="cuda", dtype=torch.float32).reshape(128, 8)
    assert torch.equal(c[:, 320:], scratch)


def make_net(cname, seed, mode, scale=1.0):
    import bdiff
    ocfg = O.config_named(cname)
    sd = O.random_state_dict(ocfg, seed, scale=scale)
    net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named(cname), mode=mode)
    net.load_state_dict(sd, strict=True)
    return net.cuda(), ocfg, sd


VARIANTS = {            # BDIFF_MEGA, BDIFF_NODE_R4
    "layers_fused": ("1", None),      # default: all layers in one persistent kernel (k_layers_tc)
    "split_r4": ("0", "1"),           # one kernel per pass, row-replicated 32-node-tile node kernel
    "split_128": ("0", "0"),          # one kernel per pass, 128-node-tile node kernel
}


def set_variant(monkeypatch, variant):
    mega, r4 = VARIANTS[variant]
    monkeypatch.setenv("BDIFF_MEGA", mega)
    if r4 is None:
        monkeypatch.delenv("BDIFF_NODE_R4", raising=False)
    else:
        monkeypatch.setenv("BDIFF_NODE_R4", r4)


@pytest.mark.parametrize("variant", list(VARIANTS))
@pytest.mark.parametrize("name", ["qm9_small_masked", "qm9_tiny_sizes", "qm9_b4_n19", "qm9_cond", "geom_mixed",
                                  "geom_max181"])
def test_tensor_forward_close_to_reference(name, variant, monkeypatch):
    set_variant(monkeypatch, variant)
    fx = load_golden(name)
    net, ocfg, sd = make_net(fx["config"], fx["weight_seed"], "tensor")
    ctx = fx["context"].cuda() if fx["context"] is not None else None
    out = net.denoise(fx["batch_index"].cuda(), fx["mask"].cuda(), fx["xh"].cuda(), fx["t"].cuda(), ctx).cpu()
    ref = fx["net_out"]
    scale = max(1.0, ref.abs().max().item())
    max_abs = (out - ref).abs().max().item()
    rms = (out - ref).pow(2).mean().sqrt().item()
    print(f"{name}: tensor-mode max|diff| {max_abs:.3e}, rms {rms:.3e} (|ref|max {ref.abs().max().item():.3g})")
    assert torch.isfinite(out).all()
    assert max_abs <= 2e-2 * scale and rms <= 5e-3 * scale


@pytest.mark.parametrize("b,variant", [(128, "layers_fused"), (128, "split_r4"), (128, "split_128"),
                                       (300, "layers_fused"), (300, "split_r4"), (300, "split_128")])
def test_tensor_and_parity_modes_agree_full_size(b, variant, monkeypatch):
    """QM9 B=128 (BASELINE config) and B=300 (more 32-node tiles than SMs): tensor mode, every kernel variant, vs parity
    mode on the same input, both on the GPU.  The fused variant goes through bdiff_profile_forward once as well, which
    fails if a dependency wait of the megakernel ever timed out."""
    set_variant(monkeypatch, variant)
    g = torch.Generator().manual_seed(4)
    nat = 19
    n = b * nat
    bi = torch.repeat_interleave(torch.arange(b), torch.full((b,), nat)).cuda()
    mask = torch.ones(n, dtype=torch.bool, device="cuda")
    xh = torch.randn((n, 9), generator=g)
    _, xc = O.centralize(xh[:, :3], bi.cpu(), mask.cpu(), b)
    xh = torch.cat((xc, xh[:, 3:]), -1).cuda()
    t = torch.full((n, 1), 0.5, device="cuda")
    outs = {}
    for mode in ("parity", "tensor"):
        net, _, _ = make_net("qm9", 7, mode)
        outs[mode] = net.denoise(bi, mask, xh, t)
        if mode == "tensor":
            prof, out2 = net.profile_forward(bi, mask, xh, t)
            assert ("layers_fused" in prof) == (variant == "layers_fused")
            # two runs differ at the bf16-rounding level: the aggregation uses floating-point atomics (any order)
            assert (out2 - outs[mode]).abs().max().item() <= 1e-2 * max(1.0, outs[mode].abs().max().item())
    d = (outs["tensor"] - outs["parity"])
    scale = max(1.0, outs["parity"].abs().max().item())
    print(f"full-size tensor vs parity: max {d.abs().max().item():.3e} rms {d.pow(2).mean().sqrt().item():.3e}")
    assert d.abs().max().item() <= 2e-2 * scale and d.pow(2).mean().sqrt().item() <= 5e-3 * scale


@pytest.mark.parametrize("name", ["chain_qm9_T6", "chain_qm9_cond_T4", "chain_geom_T3"])
def test_tensor_chain_tracks_reference_chain(name):
    """A whole sampling chain in tensor mode (CUDA-graph step, layer megakernel) against the reference's chain with the
    same recorded noise: bf16 operand rounding per forward (<= 2e-2, above) propagates through T steps, so this is a
    divergence guard, not a parity claim: 3e-2 relative on the final latent / coordinates and at least 90 % identical
    argmax atom types (measured: 3e-3..7e-3 and 97..100 %) on these 23..74-atom fixtures (parity mode meets 1e-4 and 100 % on the same fixtures,
    tests/test_gpu_parity.py); the measured values are printed (-s)."""
    import bdiff
    fx = load_golden(name)
    net, ocfg, sd = make_net(fx["config"], fx["weight_seed"], "tensor", scale=fx.get("weight_scale", 1.0))
    torch.manual_seed(fx["noise_seed"])
    sampler = bdiff.GCDMSampler(net)
    ctx = fx["context"].cuda() if fx["context"] is not None else None
    out, bi, mask, z0 = sampler.sample(torch.tensor(fx["sizes"]), ctx, num_timesteps=fx["steps"],
                                       noise=lambda s: torch.randn(s).cuda(), return_z0=True)
    rel = (z0.cpu() - fx["z_0"]).abs().max().item() / fx["z_0"].abs().max().item()
    a = ocfg.num_atom_types
    same = (out[:, 3:3 + a].cpu() == fx["out"][:, 3:3 + a]).all(dim=-1).float().mean().item()
    relx = (out[:, :3].cpu() - fx["out"][:, :3]).abs().max().item() / fx["out"][:, :3].abs().max().item()
    print(f"{name}: tensor chain z_0 rel {rel:.3e}, x rel {relx:.3e}, identical atom types {100 * same:.1f} %")
    assert rel < 3e-2 and relx < 3e-2 and same >= 0.9
