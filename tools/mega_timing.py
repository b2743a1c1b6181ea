#!/usr/bin/env python
"""Per-item timing of the layer megakernel k_layers_tc (BDIFF_TIMING=1): for a few CTAs, the first 16 work items
with their fetch time, dependency/fence wait and run time in SM cycles.  GPU only."""
import os, sys
os.environ["BDIFF_TIMING"] = "1"
ROOT = "/root/repo"
sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200")); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch, bdiff, gcpnet_oracle as O
cfg = O.config_named("qm9")
net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named("qm9"), mode="tensor")
net.load_state_dict(O.random_state_dict(cfg, 7), strict=True); net.cuda()
b, n = 128, 19
bi = torch.repeat_interleave(torch.arange(b), torch.full((b,), n)).cuda()
mask = torch.ones(b * n, dtype=torch.bool, device="cuda")
xh = torch.randn(b * n, 9, device="cuda"); t = torch.full((b * n, 1), 0.5, device="cuda")
for _ in range(3): net.denoise(bi, mask, xh, t)
torch.cuda.synchronize()
raw = net.debug_tap("dbg"); st = raw.contiguous().view(torch.int64).reshape(256, 64).cpu()
for c in (0, 1, 50, 100, 147):
    rows = st[c].reshape(16, 4).tolist()
    t0 = rows[0][1]
    out = []
    for code, tf, ts, te in rows:
        if tf == 0:
            break
        ty = "N" if (code >> 30) & 1 else "E"
        out.append(f"{ty}{(code >> 24) & 63}.{code & 0xffffff}: fetch+{tf - t0} wait {ts - tf} run {te - ts}")
    print(f"CTA {c}:\n   " + "\n   ".join(out))
