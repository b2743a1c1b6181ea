"""GPU: the CTA-pair (cta_group::2) tcgen05 self test against an fp64 matmul."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bio-diffusion_b200"))
import torch
import bdiff

lib = bdiff.load_library()
g = torch.Generator().manual_seed(0)
a = torch.randn((256, 128), generator=g).cuda()
w = torch.randn((320, 128), generator=g).cuda()
c = torch.zeros((256, 320), device="cuda")
rc = lib.bdiff_selftest_pair(C.c_void_p(torch.cuda.current_stream().cuda_stream), C.c_void_p(a.data_ptr()), C.c_void_p(w.data_ptr()),
                             C.c_void_p(c.data_ptr()))
ref = a.double() @ w.double().t()
err = (c.double() - ref).abs()
print(f"pair self test: rc={rc} rel err vs fp64 = {err.max().item() / ref.abs().max().item():.3e}; "
      f"per 128-row half {[round(err[i * 128:(i + 1) * 128].max().item() / ref.abs().max().item(), 8) for i in range(2)]}; "
      f"per 160-column half {[round(err[:, i * 160:(i + 1) * 160].max().item() / ref.abs().max().item(), 8) for i in range(2)]}")
