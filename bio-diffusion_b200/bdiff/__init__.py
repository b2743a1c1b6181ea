"""bdiff — B200-native GCPNet denoiser hot path of GCDM (bio-diffusion).

Public surface (mirrors the reference's seam):
    GCPNetDynamicsB200   drop-in for src.models.components.gcpnet.GCPNetDynamics
    GCDMSampler          inner loop of EquivariantVariationalDiffusion.mol_gen_sample
    GCDMEvalNLL          evaluation-mode NLL terms of EquivariantVariationalDiffusion.forward (forward only)
    DenoiserConfig       dims derived from the reference's Hydra config groups
"""
from .config import DenoiserConfig, parameter_shapes
from .dynamics import GCPNetDynamicsB200
from .sampler import GCDMSampler
from .loss import GCDMEvalNLL
from ._lib import BdiffError, load as load_library

__all__ = ["DenoiserConfig", "parameter_shapes", "GCPNetDynamicsB200", "GCDMSampler", "GCDMEvalNLL", "BdiffError", "load_library"]
