#!/usr/bin/env python
"""Small ncu target: a few tensor-mode denoiser forwards on the BASELINE workload (QM9 B=128 x 19 atoms, or
`geom`: B=64 x 44), nothing else (no chain, no CPU baseline), so `ncu --set full -k regex:... -s N -c 1` is quick.
  ncu ... python tools/ncu_target.py [qm9|geom] [n_forwards]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import bdiff  # noqa: E402
import gcpnet_oracle as O  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "qm9"
nfwd = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cfg = O.config_named(name)
net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named(name), mode="tensor")
net.load_state_dict(O.random_state_dict(cfg, 7), strict=True)
net.cuda()
b, n = (128, 19) if name.startswith("qm9") else (64, 44)
bi = torch.repeat_interleave(torch.arange(b), torch.full((b,), n)).cuda()
mask = torch.ones(b * n, dtype=torch.bool, device="cuda")
g = torch.Generator().manual_seed(1)
xh = torch.randn((b * n, 3 + cfg.num_h), generator=g).cuda()
t = torch.full((b * n, 1), 0.5, device="cuda")
for _ in range(nfwd):
    out = net.denoise(bi, mask, xh, t)
torch.cuda.synchronize()
print("ok", tuple(out.shape), float(out.abs().max()))
