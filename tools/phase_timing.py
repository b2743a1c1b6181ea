#!/usr/bin/env python
"""Dump the in-kernel clock64 phase stamps of the tensor-core kernels (BDIFF_TIMING=1).  GPU only.
Runs one tensor-mode forward on the QM9 B=128 workload and prints, for a few CTAs, the cycle deltas between the
stamps the LAST launched TC kernel of each kind left in the debug buffer (node kernel = last writer)."""
import os
import sys

os.environ["BDIFF_TIMING"] = "1"
os.environ.setdefault("BDIFF_MEGA", "0")      # per-pass kernels: their per-phase stamps (tools/mega_timing.py: per-item)
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import bdiff  # noqa: E402
import gcpnet_oracle as O  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "node"
cfg = O.config_named("qm9")
net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named("qm9"), mode="tensor")
net.load_state_dict(O.random_state_dict(cfg, 7), strict=True)
net.cuda()
b, n = 128, 19
bi = torch.repeat_interleave(torch.arange(b), torch.full((b,), n)).cuda()
mask = torch.ones(b * n, dtype=torch.bool, device="cuda")
xh = torch.randn(b * n, 9, device="cuda")
t = torch.full((b * n, 1), 0.5, device="cuda")
for _ in range(3):
    net.denoise(bi, mask, xh, t)
torch.cuda.synchronize()
raw = net.debug_tap("dbg")                      # [256, 128] float32 view of [256][64] int64
st = raw.contiguous().view(torch.int64).reshape(256, 64).cpu()
# the node kernel ran last (<= 76 CTAs at this size); the edge kernel stamps survive in CTAs above that
rows = range(0, 4) if which == "node" else range(100, 104)
for c in rows:
    fine = [v for v in st[c, 52:64].tolist() if v > 0]
    e = st[c, :28].tolist()
    m = st[c, 28:52].tolist()
    e = [v for v in e if v > 0]
    m = [v for v in m if v > 0]
    if not e:
        continue
    t0 = e[0]
    print(f"CTA {c}: epilogue stamps (cycles since tile start):", [v - t0 for v in e])
    print(f"        deltas:", [e[i + 1] - e[i] for i in range(len(e) - 1)])
    if fine:
        print(f"        fine stamps (T0 sub-phases):", [v - t0 for v in fine])
    if m:
        print(f"        MMA-thread stamps:", [v - t0 for v in m])
