#!/usr/bin/env python
"""Times bdiff_optimizer_step (clip + AdamW(amsgrad) + EMA, 3 kernels) on the full QM9 denoiser parameter set and on a
1 GiB synthetic set (larger than L2), against torch.optim.AdamW(foreach) + clip_grad_norm_ + a foreach EMA on the same
GPU.  Prints one JSON line.  Algorithmic bytes per element: 4 (norm pass) + 24 read + 20 written = 48."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "bio-diffusion_b200"))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch  # noqa: E402
import bdiff  # noqa: E402
from bdiff.optim import GCDMTrainTail  # noqa: E402
import gcpnet_oracle as O  # noqa: E402


def timeit(fn, n=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / n


def run(params, label):
    n = sum(p.numel() for p in params)
    opt = GCDMTrainTail(params)
    for p in params:
        p.grad.normal_()
    ms = timeit(opt.step)
    ref_params = [torch.nn.Parameter(p.detach().clone()) for p in params]
    for q in ref_params:
        q.grad = torch.randn_like(q)
    ref = torch.optim.AdamW(ref_params, lr=1e-4, weight_decay=1e-12, amsgrad=True, foreach=True)
    ema = [q.detach().clone() for q in ref_params]

    def ref_step():
        torch.nn.utils.clip_grad_norm_(ref_params, 4500.0)
        ref.step()
        torch._foreach_mul_(ema, 0.9999)
        torch._foreach_add_(ema, [q.data for q in ref_params], alpha=1e-4)
    ms_ref = timeit(ref_step)
    return {"set": label, "tensors": len(params), "elements": n, "ms": ms, "gbs": 48.0 * n / ms / 1e6,
            "torch_foreach_ms": ms_ref}


net = bdiff.GCPNetDynamicsB200(config=bdiff.DenoiserConfig.named("qm9"), mode="parity")
net.load_state_dict(O.random_state_dict(O.config_named("qm9"), 3), strict=True)
net.cuda()
out = [run(list(net.parameters()), "qm9 denoiser (432 tensors)")]
big = [torch.nn.Parameter(torch.randn(1 << 24, device="cuda")) for _ in range(3)]      # 3 x 64 MiB x 6 arrays > L2
out.append(run(big, "synthetic 3 x 16M"))
peaks = {}
try:
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
except Exception:
    pass
print(json.dumps({"optimizer_tail": out, "hbm_peak_gbs": peaks.get("hbm_gbs")}))
