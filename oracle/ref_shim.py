"""Import shim that lets the UNMODIFIED reference modules run in the build container.

TEST INFRASTRUCTURE ONLY.  This file is used by `tests/golden/make_golden.py` (fixture
generation) and by `oracle/check_against_reference.py` (validation of the CPU restatement in
`oracle/gcpnet_oracle.py`).  It needs `/root/reference`, which does not exist on the GPU box, so
nothing under `tests/ -m gpu`, `bench.py` or `__graft_entry__.smoke()` may import it.

What it does: pre-populates `sys.modules` with stand-ins for the third-party packages the
reference imports but this image lacks:

  functional stubs (arithmetic the hot path really executes)
    torch_scatter.scatter          -> index_add_ based sum / mean (pytorch-scatter 2.1.0 semantics:
                                      dim=0, dim_size honoured, mean divides by max(count, 1))
    torch_geometric.data.Batch     -> attribute bag with __getitem__/__setitem__/num_nodes
    omegaconf.DictConfig/OmegaConf -> attribute dict with __copy__, to_container, open_dict
  pass-through stubs
    torchtyping.TensorType / patch_typeguard, typeguard.typechecked
  MagicMock modules for everything else (lightning, rdkit, matplotlib, wandb, hydra, ...).

No reference source is copied; the reference is imported from where it lies.
"""
from __future__ import annotations

import sys
import types
from unittest.mock import MagicMock

import torch

REFERENCE_ROOT = "/root/reference"


# ----------------------------------------------------------------------------- functional stubs
def _scatter(src, index, dim=0, out=None, dim_size=None, reduce="sum"):
    assert dim == 0, "the hot path only scatters along dim 0"
    if dim_size is None:
        dim_size = int(index.max().item()) + 1 if index.numel() else 0
    shape = (dim_size,) + tuple(src.shape[1:])
    res = torch.zeros(shape, dtype=src.dtype, device=src.device)
    res.index_add_(0, index, src)
    if reduce in ("sum", "add"):
        return res
    if reduce == "mean":
        cnt = torch.zeros(dim_size, dtype=src.dtype, device=src.device)
        cnt.index_add_(0, index, torch.ones_like(index, dtype=src.dtype))
        cnt = cnt.clamp(min=1).view((dim_size,) + (1,) * (src.dim() - 1))
        return res / cnt
    raise NotImplementedError(reduce)


def _unbatch(src, batch, dim=0):
    """torch_geometric.utils.unbatch (PyG 2.2.0): split `src` along `dim` by the (sorted) batch vector."""
    sizes = torch.bincount(batch).tolist()
    return src.split(sizes, dim)


class _Batch:
    """Attribute bag standing in for torch_geometric.data.Batch on the sampling path."""

    def __init__(self, **kw):
        for k, v in kw.items():
            setattr(self, k, v)

    def __getitem__(self, key):
        return getattr(self, key)

    def __setitem__(self, key, value):
        setattr(self, key, value)

    def __contains__(self, key):
        return hasattr(self, key)

    @property
    def num_nodes(self):
        for k in ("batch", "x", "mask", "h"):
            v = getattr(self, k, None)
            if isinstance(v, torch.Tensor):
                return v.shape[0]
        raise AttributeError("num_nodes")


class _DictConfig(dict):
    """Attribute dict standing in for omegaconf.DictConfig."""

    def __getattr__(self, k):
        try:
            return self[k]
        except KeyError as e:
            raise AttributeError(k) from e

    def __setattr__(self, k, v):
        self[k] = v

    def __copy__(self):
        return _DictConfig(self)

    def __deepcopy__(self, memo):
        import copy as _c
        return _DictConfig({k: _c.deepcopy(v, memo) for k, v in self.items()})


class _OmegaConf:
    @staticmethod
    def to_container(cfg, **kw):
        return dict(cfg)

    @staticmethod
    def create(d=None):
        return _DictConfig(d or {})


class _open_dict:
    def __init__(self, cfg):
        self.cfg = cfg

    def __enter__(self):
        return self.cfg

    def __exit__(self, *a):
        return False


class _TensorTypeMeta(type):
    def __getitem__(cls, item):
        return torch.Tensor


class _TensorType(metaclass=_TensorTypeMeta):
    pass


def _typechecked(f=None, **kw):
    if f is None:
        return lambda g: g
    return f


def _module(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    return m


_INSTALLED = False


def install():
    """Install the stubs and put the reference on sys.path (idempotent)."""
    global _INSTALLED
    if _INSTALLED:
        return
    import os
    if not os.path.isdir(REFERENCE_ROOT):
        raise RuntimeError(f"{REFERENCE_ROOT} not present: the reference shim only works in the build container")

    sm = sys.modules
    sm["torch_scatter"] = _module("torch_scatter", scatter=_scatter)

    tg = _module("torch_geometric")
    tgd = _module("torch_geometric.data", Batch=_Batch, Data=_Batch, Dataset=object)
    tg.data = tgd
    sm["torch_geometric"] = tg
    sm["torch_geometric.data"] = tgd
    for sub in ("loader", "nn", "transforms"):
        sm[f"torch_geometric.{sub}"] = MagicMock()
    sm["torch_geometric.utils"] = _module("torch_geometric.utils", unbatch=_unbatch)

    sm["omegaconf"] = _module("omegaconf", DictConfig=_DictConfig, OmegaConf=_OmegaConf,
                              open_dict=_open_dict, ListConfig=list)
    sm["torchtyping"] = _module("torchtyping", TensorType=_TensorType, patch_typeguard=lambda: None)
    sm["typeguard"] = _module("typeguard", typechecked=_typechecked)

    # pytorch_lightning: Callback / ModelCheckpoint are subclassed by src/utils/__init__.py
    pl = MagicMock()
    pl.Callback = type("Callback", (), {})
    plc = MagicMock()
    plc.ModelCheckpoint = type("ModelCheckpoint", (), {})
    plu = MagicMock()
    plu.rank_zero_only = lambda f: f
    pl.utilities = plu
    pl.callbacks = plc
    sm["pytorch_lightning"] = pl
    sm["pytorch_lightning.callbacks"] = plc
    sm["pytorch_lightning.utilities"] = plu
    for sub in ("utilities.exceptions", "utilities.types", "loggers", "loggers.logger", "core",
                "utilities.rank_zero", "utilities.memory"):
        sm[f"pytorch_lightning.{sub}"] = MagicMock()

    for name in ("torch_cluster", "hydra", "hydra.utils", "hydra.core", "hydra.core.hydra_config",
                 "torchmetrics", "rdkit", "rdkit.Chem", "prody", "matplotlib", "matplotlib.pyplot",
                 "matplotlib.lines", "matplotlib.axes", "matplotlib.axes._subplots", "imageio",
                 "torchviz", "pyrootutils", "wandb", "wandb.sdk", "wandb.sdk.wandb_run", "rich",
                 "rich.prompt", "rich.syntax", "rich.tree", "pymol", "Bio", "Bio.PDB", "scipy.spatial",
                 "msgpack", "tqdm", "sklearn", "networkx", "posebusters"):
        if name in ("tqdm", "msgpack"):
            continue
        sm[name] = MagicMock()

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    _INSTALLED = True


def qm9_cfgs(conditioning=(), include_charges=True, num_atom_types=5, geom=False):
    """The four Hydra config groups the denoiser constructor reads, as plain attr-dicts.

    Values restate configs/model/{model_cfg,module_cfg,layer_cfg,diffusion_cfg}/*.yaml and
    configs/datamodule/dataloader_cfg/edm_{qm9,geom}_dataloader.yaml.
    """
    install()
    from src.models.components.gcpnet import GCP2
    from functools import partial
    model_cfg = _DictConfig(
        chi_input_dim=2, e_input_dim=1, xi_input_dim=1,
        h_hidden_dim=256, chi_hidden_dim=32,
        e_hidden_dim=16 if geom else 64, xi_hidden_dim=8 if geom else 16,
        num_encoder_layers=4 if geom else 9, num_decoder_layers=3, dropout=0.0)
    module_cfg = _DictConfig(
        selected_GCP=partial(GCP2), norm_x_diff=True, scalar_gate=0, vector_gate=True,
        vector_residual=False, vector_frame_residual=False, frame_gate=False, sigma_frame_gate=False,
        scalar_nonlinearity="silu", vector_nonlinearity="silu", nonlinearities=["silu", "silu"],
        bottleneck=4, vector_linear=True, vector_identity=True, default_vector_residual=False,
        default_bottleneck=4, node_positions_weight=1.0, update_positions_with_vector_sum=False,
        ablate_frame_updates=False, ablate_scalars=False, ablate_vectors=False,
        conditioning=list(conditioning), clip_gradients=True, log_grad_flow_steps=500)
    mp_cfg = _DictConfig(edge_encoder=False, edge_gate=False, num_message_layers=4, message_residual=0,
                         message_ff_multiplier=1, self_message=True, use_residual_message_gcp=True)
    layer_cfg = _DictConfig(mp_cfg=mp_cfg, pre_norm=False, use_gcp_norm=False, use_gcp_dropout=False,
                            use_scalar_message_attention=True, num_feedforward_layers=1, dropout=0.0,
                            nonlinearity_slope=1e-2)
    diffusion_cfg = _DictConfig(
        ddpm_mode="unconditional", dynamics_network="gcpnet", diffusion_target="atom_types_and_coords",
        num_timesteps=1000, parametrization="eps", noise_schedule="polynomial_2", noise_precision=1e-5,
        loss_type="l2", norm_values=[1.0, 4.0, 10.0] if not conditioning else [1.0, 8.0, 1.0],
        norm_biases=[None, 0.0, 0.0], condition_on_time=True, self_condition=False,
        norm_training_by_max_nodes=False)
    if geom:
        diffusion_cfg["norm_values"] = [1.0, 4.0, 10.0]
    dataloader_cfg = _DictConfig(num_atom_types=num_atom_types, include_charges=include_charges,
                                 num_x_dims=3, num_radials=1, remove_h=False)
    return model_cfg, module_cfg, layer_cfg, diffusion_cfg, dataloader_cfg


def build_reference_dynamics(config="qm9", seed=0):
    """Instantiate the reference GCPNetDynamics (default init under torch.manual_seed(seed))."""
    install()
    from src.models.components.gcpnet import GCPNetDynamics
    if config == "qm9":
        cfgs = qm9_cfgs()
    elif config == "qm9_cond":
        cfgs = qm9_cfgs(conditioning=("alpha",), include_charges=False)
    elif config == "geom":
        cfgs = qm9_cfgs(geom=True, include_charges=False, num_atom_types=16)
    else:
        raise ValueError(config)
    torch.manual_seed(seed)
    net = GCPNetDynamics(*cfgs)
    net.eval()
    return net, cfgs


def build_reference_ddpm(config="qm9", seed=0, n_nodes_hist=None):
    """Reference EquivariantVariationalDiffusion wrapped around the reference denoiser."""
    install()
    from src.models.components.variational_diffusion import EquivariantVariationalDiffusion
    net, cfgs = build_reference_dynamics(config, seed)
    _, _, _, diffusion_cfg, dataloader_cfg = cfgs
    hist = n_nodes_hist or {19: 1}
    dcfg = _DictConfig(diffusion_cfg)
    dcfg["verbose"] = False
    ddpm = EquivariantVariationalDiffusion(
        dynamics_network=net, diffusion_cfg=dcfg, dataloader_cfg=dataloader_cfg,
        dataset_info={"n_nodes": hist})
    ddpm.eval()
    return ddpm, cfgs


Batch = _Batch
