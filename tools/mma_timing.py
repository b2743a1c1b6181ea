"""GPU: cycles per tcgen05.mma (cta_group::1, M=128, K=16, SS operands) for the instruction mixes the megakernel issues."""
import ctypes as C
import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "bio-diffusion_b200"))
import torch
import bdiff

lib = bdiff.load_library()
a = torch.randn((128, 128)).cuda()
w = torch.randn((320, 128)).cuda()
names = {16: "N=256 x96", 17: "N=32 x96", 18: "N=64 x96", 19: "N=160 x96", 20: "(256|32|32) x3 x16 steps", 21: "(256|64) x3 x16 steps",
         22: "(160|160) x3 x16 steps", 23: "N=256 x4 x16 steps (node)", 24: "(160|160)x3 + wait,commit per plane",
         25: "(160|160)x3 + waits, commit per K step", 26: "(160|160)x3 + commit per plane"}
for variant in sorted(names):
    best = None
    for rep in range(3):
        c = torch.zeros((128, 336), device="cuda")
        rc = lib.bdiff_selftest_split(C.c_void_p(torch.cuda.current_stream().cuda_stream), variant, C.c_void_p(a.data_ptr()),
                                      C.c_void_p(w.data_ptr()), C.c_void_p(c.data_ptr()))
        assert rc == 0
        cyc, n, issue = c[0, 0].item(), c[0, 1].item(), c[0, 2].item()
        best = (cyc, n, issue) if best is None or cyc < best[0] else best
    cyc, n, issue = best
    print(f"variant {variant} {names[variant]:28s}: {cyc:8.0f} cycles for {n:4.0f} MMAs = {cyc / n:6.1f} cyc/MMA (issue loop alone {issue:6.0f} = {issue / n:5.1f}/MMA)")
