"""CPU: the C-ABI library loads, exports every symbol include/bdiff.h declares, and fails loudly without a GPU.
No compute calls are made here."""
import ctypes as C
import os
import re

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "bdiff.h")).read()
    return sorted(set(re.findall(r"BDIFF_API\s+[\w\s\*]+?\b(bdiff_\w+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    import bdiff
    from bdiff import _lib
    names = declared_symbols()
    assert len(names) >= 15
    lib = bdiff.load_library()
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/bdiff.h but not exported"
        assert n in _lib.PROTOTYPES, f"{n} has no ctypes prototype"
    assert set(_lib.PROTOTYPES) == set(names)
    assert lib.bdiff_abi_version() == 1


def test_create_rejects_bad_configs_and_missing_gpu():
    from bdiff import _lib
    lib = _lib.load()
    h = C.c_void_p()
    bad = _lib.Config(num_h=6, num_context=0, num_layers=9, h_hidden=128, chi_hidden=32, e_hidden=64, xi_hidden=16, mode=0)
    assert lib.bdiff_create(C.byref(bad), C.byref(h)) == -1
    assert b"h_hidden" in lib.bdiff_last_error(None)
    if not torch.cuda.is_available():
        ok = _lib.Config(num_h=6, num_context=0, num_layers=9, h_hidden=256, chi_hidden=32, e_hidden=64, xi_hidden=16, mode=0)
        assert lib.bdiff_create(C.byref(ok), C.byref(h)) == -2
        assert b"no CPU fallback" in lib.bdiff_last_error(None)


def test_module_matches_reference_checkpoint_contract():
    """Parameter names / shapes / counts of the three shipped configs."""
    import bdiff
    import gcpnet_oracle as O
    expect = {"qm9": (6213433, 432), "qm9_cond": (6213433, 432), "geom": (2727387, 202)}
    for name, (nparam, ntens) in expect.items():
        net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named(name))
        sd = net.state_dict()
        assert sum(v.numel() for v in sd.values()) == nparam and len(sd) == ntens
        ref_shapes = O.param_shapes(O.config_named(name))
        assert {k: tuple(v.shape) for k, v in sd.items()} == ref_shapes
        net.load_state_dict(O.random_state_dict(O.config_named(name), 1), strict=True)


def test_no_cpu_fallback_on_forward():
    import bdiff
    net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named("qm9"))
    bi = torch.zeros(3, dtype=torch.long)
    with pytest.raises(bdiff.BdiffError):
        net.denoise(bi, torch.ones(3, dtype=torch.bool), torch.zeros(3, 9), torch.zeros(3, 1))


def test_config_from_reference_cfgs_and_rejections():
    import bdiff
    model = dict(chi_input_dim=2, e_input_dim=1, xi_input_dim=1, h_hidden_dim=256, chi_hidden_dim=32, e_hidden_dim=16,
                 xi_hidden_dim=8, num_encoder_layers=4, dropout=0.0)
    module = dict(norm_x_diff=True, scalar_gate=0, vector_gate=True, frame_gate=False, nonlinearities=["silu", "silu"],
                  bottleneck=4, default_bottleneck=4, conditioning=[], vector_residual=False)
    layer = dict(pre_norm=False, use_gcp_norm=False, use_gcp_dropout=False, use_scalar_message_attention=True,
                 num_feedforward_layers=1, mp_cfg=dict(num_message_layers=4, use_residual_message_gcp=True))
    diff = dict(diffusion_target="atom_types_and_coords", self_condition=False, condition_on_time=True,
                num_timesteps=1000, noise_precision=1e-5, noise_schedule="polynomial_2", norm_values=[1.0, 4.0, 10.0],
                norm_biases=[None, 0.0, 0.0])
    data = dict(num_atom_types=16, include_charges=False, num_x_dims=3)
    cfg = bdiff.DenoiserConfig.from_reference_cfgs(model, module, layer, diff, data)
    assert cfg == bdiff.DenoiserConfig.named("geom")
    with pytest.raises(NotImplementedError):
        bdiff.DenoiserConfig.from_reference_cfgs(model, dict(module, frame_gate=True), layer, diff, data)
    with pytest.raises(NotImplementedError):
        bdiff.DenoiserConfig.from_reference_cfgs(model, module, layer, dict(diff, self_condition=True), data)


def test_schedule_matches_oracle():
    import gcpnet_oracle as O
    from bdiff.schedule import decode_coefficients, gamma_table, step_coefficient_table
    g = gamma_table()
    assert torch.equal(g, O.gamma_table())
    tab = step_coefficient_table(g, 50)
    for r, s in enumerate(reversed(range(50))):
        s_t = torch.tensor(s / 50, dtype=torch.float32)
        t_t = torch.tensor((s + 1) / 50, dtype=torch.float32)
        a, c, sig = O.step_coefficients(g[torch.round(s_t * 1000).long()], g[torch.round(t_t * 1000).long()])
        # vectorised vs scalar softplus/logsigmoid may differ in the last bit
        assert torch.allclose(tab[r], torch.stack((a, c, sig, t_t)), rtol=2e-6, atol=1e-7)
    d = decode_coefficients(g)
    assert torch.isclose(d[2], torch.exp(0.5 * g[0]))
