"""GPU: does the NaN guard (gcpnet.py:1214-1216) of an UNTRAINED GEOM-Drugs denoiser fire in the reference-precision (parity,
all-fp32 FFMA) chain as well?  64 molecules with sizes ~ dataset histogram, T=1000, same device noise stream in both modes."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, bdiff, gcpnet_oracle as O
from bdiff.datasets import GEOM_N_NODES, sample_num_nodes

sizes = sample_num_nodes(GEOM_N_NODES, 512, seed=123)[:64]
for mode in ("parity", "tensor"):
    cfg = O.config_named("geom")
    net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named("geom"), mode=mode)
    net.load_state_dict(O.random_state_dict(cfg, 7), strict=True); net.cuda()
    s = bdiff.GCDMSampler(net)
    torch.manual_seed(123)
    out, bi, mask = s.sample(sizes, num_timesteps=1000, record_moments=True)
    m = s.last_moments.cpu()
    first_bad = (~torch.isfinite(m).all(dim=1)).nonzero()
    print(f"{mode}: nan_guard_hits {s.nan_guard_count()}, finite out {bool(torch.isfinite(out).all())}, "
          f"first non-finite moment step {first_bad[0].item() if len(first_bad) else None}, moments at steps 100/500/900/999: "
          f"{[m[i].tolist() for i in (100, 500, 900, 999)]}")
