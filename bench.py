#!/usr/bin/env python
"""bench.py — molecules/s for 1000-step GCDM sampling with the B200-native GCPNet denoiser.

Contract:
  python bench.py --gpus N --steps K --warmup W            # our arm (torchrun for N > 1, one rank per GPU)
  python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path (oracle port) on host cores

One bench "step" = one complete sample of the batch: T reverse-diffusion steps + the final decode
(T+1 denoiser forwards), i.e. BASELINE.json's metric "molecules/sec (1000-step sample)".  Workload at N=1 is
BASELINE config[1]: QM9 unconditional, T=1000, batch 128 (19 atoms per molecule, the README demo size);
for N > 1 every GPU gets its own 128 molecules (weak scaling) and the final coordinates are all-gathered once.
Rank 0 prints ONE JSON line.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "bio-diffusion_b200"), os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

import torch  # noqa: E402

METRIC = "molecules/sec (1000-step sample)"


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="qm9", choices=["qm9", "qm9_cond", "geom"])
    ap.add_argument("--batch", type=int, default=None, help="molecules per GPU (default 128; geom 64)")
    ap.add_argument("--atoms", type=int, default=None, help="atoms per molecule (default 19 qm9 / 44 geom)")
    ap.add_argument("--timesteps", type=int, default=1000)
    ap.add_argument("--mode", default=os.environ.get("BDIFF_MODE", "tensor"), choices=["parity", "tensor"],
                    help="tensor: tcgen05 bf16-operand GEMMs with fp32 accumulation (default); parity: all-fp32 FFMA")
    ap.add_argument("--no-parity-leg", action="store_true", help="skip the extra fp32 parity-mode chain (tensor mode)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ clocks
class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index=0):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None
        self.th = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200", "-i", str(self.gpu)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        self.th = threading.Thread(target=self._read, daemon=True)
        self.th.start()

    def _read(self):
        for line in self.proc.stdout:
            parts = [x.strip() for x in line.split(",")]
            if len(parts) >= 9:
                self.rows.append(parts)

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, pw, reasons = [], [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])); mx.append(float(r[2])); pw.append(float(r[3]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(pw) if pw else None, "samples": len(sm), "reasons": sorted(reasons)}


# --------------------------------------------------------------------------------------------- workloads
def workload(args):
    batch = args.batch or (64 if args.config == "geom" else 128)
    atoms = args.atoms or (44 if args.config == "geom" else 19)
    return batch, atoms


def cpu_reference_chain(config, batch, atoms, steps, seed=123):
    """The reference's CPU path (oracle port of the PyG/torch_scatter code) for one bounded chain.
    Returns seconds.  Only used as the reported baseline / reference arm."""
    import gcpnet_oracle as O
    ocfg = O.config_named(config)
    sd = O.random_state_dict(ocfg, seed=7)
    num_nodes = torch.full((batch,), atoms, dtype=torch.long)
    ctx = torch.randn((batch, ocfg.num_context), generator=torch.Generator().manual_seed(seed)) if ocfg.num_context else None
    noise = O.SeededNoise(seed)
    t0 = time.perf_counter()
    with torch.no_grad():
        O.sample_chain(sd, ocfg, num_nodes, noise, num_timesteps=steps, context=ctx)
    return time.perf_counter() - t0


def pick_cpu_threads(config, atoms):
    """All the host threads the CPU path can USE: tiny tensors get slower when over-threaded, so try a few
    intra-op thread counts on a 2-step chain and keep the fastest (the count is reported as `cores`)."""
    cores = os.cpu_count() or 1
    best, best_t = 1, None
    for c in sorted({min(cores, x) for x in (4, 8, 16, 32, cores)}):
        torch.set_num_threads(c)
        cpu_reference_chain(config, 2, atoms, 1)
        t = cpu_reference_chain(config, 4, atoms, 2)
        if best_t is None or t < best_t:
            best, best_t = c, t
    torch.set_num_threads(best)
    return best


def cpu_baseline_block(args, reps=1):
    """config[0]: QM9 unconditional, batch 4, T=50 on the host cores; extrapolated linearly to T=1000."""
    b, n, t = 4, 19 if args.config != "geom" else 44, 50
    pick_cpu_threads(args.config, n)
    secs = min(cpu_reference_chain(args.config, b, n, t) for _ in range(reps))
    fwd = t + 1
    value = b / (secs * (args.timesteps + 1) / fwd)
    return {"value": value, "unit": "molecules/s", "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"oracle port of the reference PyG path: batch {b} x {n} atoms, T={t} ({fwd} denoiser forwards) "
                      f"in {secs:.2f} s = {1000 * secs / fwd:.1f} ms/forward; extrapolated linearly to T={args.timesteps}",
            "seconds": secs, "ms_per_forward": 1000 * secs / fwd}


def run_reference(args):
    """--impl reference: the reference's CPU implementation (oracle port; /root/reference does not travel to the
    GPU box and needs PyG/torch_scatter which are not installable offline) on all host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    b, n, t = 4, 19 if args.config != "geom" else 44, 50
    pick_cpu_threads(args.config, n)
    for _ in range(max(1, min(args.warmup, 2))):
        cpu_reference_chain(args.config, 2, n, 2)
    times = [cpu_reference_chain(args.config, b, n, t) for _ in range(args.steps)]
    secs = sum(times) / len(times)
    fwd = t + 1
    value = b / (secs * (args.timesteps + 1) / fwd)
    batch, atoms = workload(args)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000 * secs, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{args.config} sampling, T={args.timesteps}, batch {batch} x {atoms} atoms per GPU",
                   "sample": f"each step = batch {b} x {n} atoms, T={t} on host cores, extrapolated to T={args.timesteps}"},
        "cpu_baseline": {"value": value, "unit": "molecules/s", "cores": torch.get_num_threads(), "kind": "port",
                         "sample": f"batch {b} x {n} atoms, T={t}, {1000 * secs / fwd:.1f} ms/forward"},
        "e2e": {"value": value, "unit": "molecules/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ our arm
def run_ours(args):
    import torch.distributed as dist
    import bdiff
    import gcpnet_oracle as O   # only for seeded synthetic weights (shapes/magnitudes), not on the timed path

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (our arm) needs a CUDA device; there is no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ["NCCL_DEBUG"] = "WARN"        # keep stdout to the single JSON line (no "NCCL version ..." banner)
        dist.init_process_group("nccl", device_id=dev)

    batch, atoms = workload(args)
    T = args.timesteps
    dcfg = bdiff.DenoiserConfig.named(args.config)
    ocfg = O.config_named(args.config)
    sd = O.random_state_dict(ocfg, seed=7)       # synthetic random-init weights of the named architecture
    net = bdiff.GCPNetDynamicsB200(config=dcfg, mode=args.mode)
    net.load_state_dict(sd, strict=True)
    net.to(dev)
    sampler = bdiff.GCDMSampler(net, use_cuda_graph=True)
    torch.manual_seed(123 + rank)

    num_nodes_host = torch.full((batch,), atoms, dtype=torch.long).pin_memory()
    num_nodes_dev = num_nodes_host.to(dev)
    ctx_host = ctx_dev = None
    if dcfg.num_context:
        ctx_host = torch.randn((batch, dcfg.num_context), generator=torch.Generator().manual_seed(5)).pin_memory()
        ctx_dev = ctx_host.to(dev)
    n_nodes = batch * atoms
    out_host = torch.empty((n_nodes, 3 + dcfg.num_atom_types + int(dcfg.include_charges)), pin_memory=True)
    flush_buf = torch.empty(256 * 1024 * 1024 // 4, device=dev)      # > 126 MB L2

    def gather(out):
        if world > 1:
            bufs = [torch.empty_like(out) for _ in range(world)]
            dist.all_gather(bufs, out)                                # single NCCL gather of final coordinates
            return bufs
        return [out]

    def chain_resident():
        out, _, _ = sampler.sample(num_nodes_dev, ctx_dev, T)
        gather(out)

    def chain_e2e():
        nn_dev = num_nodes_host.to(dev, non_blocking=True)            # H2D of this step's inputs (pinned)
        cdev = ctx_host.to(dev, non_blocking=True) if ctx_host is not None else None
        out, _, _ = sampler.sample(nn_dev, cdev, T)
        gather(out)
        out_host.copy_(out, non_blocking=True)                        # D2H of the step's result
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        total = 0.0
        barrier()
        for _ in range(k):
            flush_buf.fill_(1.0)                                      # L2 flush between timed iterations (untimed)
            torch.cuda.synchronize()
            ev0.record()
            fn()
            ev1.record()
            torch.cuda.synchronize()
            total += ev0.elapsed_time(ev1)
        barrier()
        t = torch.tensor([total], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)                  # max over ranks
        return t.item() / 1000.0

    def log(msg):
        if rank == 0:
            print(f"[bench +{time.perf_counter() - t_start:7.1f}s] {msg}", file=sys.stderr, flush=True)

    t_start = time.perf_counter()
    for i in range(args.warmup):
        chain_resident()
        torch.cuda.synchronize()
        log(f"warm-up chain {i + 1}/{args.warmup} done")
    clocks = ClockSampler(local)
    launches0 = sampler_launches(sampler, net)
    if rank == 0:
        clocks.start()
    secs = timed(chain_resident, args.steps)
    clk = clocks.stop() if rank == 0 else None
    launches = sampler_launches(sampler, net) - launches0
    log(f"timed resident chains done: {secs:.2f} s for {args.steps}")
    secs_e2e = timed(chain_e2e, args.steps)
    log(f"timed e2e chains done: {secs_e2e:.2f} s")

    mols_total = batch * world * args.steps
    value = mols_total / secs
    e2e_value = mols_total / secs_e2e

    # ---- roofline of the dominant kernel (fused message + scatter), timed live with CUDA events in the library
    bi = torch.repeat_interleave(torch.arange(batch, device=dev), num_nodes_dev)
    mask = torch.ones(n_nodes, dtype=torch.bool, device=dev)
    g = torch.Generator().manual_seed(3)
    xh = torch.randn((n_nodes, 3 + dcfg.num_h), generator=g).to(dev)
    tt = torch.full((n_nodes, 1), 0.5, device=dev)
    cnode = ctx_dev[bi] if ctx_dev is not None else None
    prof = None
    for i in range(6):
        flush_buf.fill_(0.0) if i else None
        pr, _ = net.profile_forward(bi, mask, xh, tt, cnode, batch)
        if i:   # first call is warm-up
            prof = pr if prof is None else {k: prof[k] + pr[k] for k in pr}
    prof = {k: v / 5 for k, v in prof.items()}
    E = batch * atoms * atoms
    L = dcfg.num_layers
    ed, xd = dcfg.e_hidden, dcfg.xi_hidden
    hid0 = (64 + xd) // 4
    w_msg = (256 * (512 + ed + hid0 + 9) + 256 + hid0 * (64 + xd) + 3 * (64 + xd) + 32 * hid0 + 32 * 256 + 32
             + 3 * (256 * 273 + 256 + 8 * 32 + 3 * 32 + 32 * 8 + 32 * 256 + 32) + 257)
    bytes_alg = E * (4 * (ed + 3 * xd) + 36) + n_nodes * (2 * 4 * 352) + 4 * w_msg
    flops_edge = 821176 if args.config != "geom" else 793224                              # per edge per layer
    fused = "layers_fused" in prof        # tensor mode default: one persistent kernel runs all L edge + node passes
    flops_node = 575324                                                                   # per node per layer
    if fused:
        t_kernel = prof["layers_fused"] / 1000.0
        launch_flops = L * (E * flops_edge + n_nodes * flops_node)
        bytes_alg = L * bytes_alg
    else:
        t_kernel = prof["edge_message"] / L / 1000.0
        launch_flops = E * flops_edge
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    traffic = None
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "roofline_traffic.json"))).get("tensor_fused" if fused else args.mode)
    except Exception:
        pass
    achieved_gbs = bytes_alg / t_kernel / 1e9
    achieved_tf = launch_flops / t_kernel / 1e12
    tensor_peak = float(peaks.get("bf16_tflops_sustained", 1400.0))     # the kernel is timed inside a long step
    kname = ("k_layers_tc (persistent tcgen05 kernel: the fused per-edge message MLP + segmented scatter-sum and the node "
             "update of all %d layers, tiles scheduled by dependency flags)" % L if fused
             else "k_edge_message_tc (tcgen05 fused per-edge GCP message MLP + segmented scatter-sum)" if args.mode == "tensor"
             else "k_edge_message (fp32 fused per-edge GCP message MLP + segmented scatter-sum)")
    common = {
        "kernel": kname, "traffic": traffic, "algorithmic_bytes_per_launch": bytes_alg,
        "algorithmic_flops_per_launch": launch_flops, "kernel_ms": t_kernel * 1000,
        "hbm_achieved_gbs": achieved_gbs, "hbm_peak_gbs": hbm_peak, "hbm_frac": achieved_gbs / hbm_peak,
        "algorithmic_tflops": achieved_tf,
        "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback (B200_PROFILING.md)",
        "note": "the fused pass is compute-bound by construction (~1.3 kFLOP/B, SURVEY.md fact 3); the HBM figure "
                "(BASELINE.json's metric) is carried as hbm_* next to the binding roof",
        "forward_ms_by_kernel": prof,
    }
    if args.mode == "tensor":
        roofline = dict(bound="tensor", achieved=achieved_tf, peak=tensor_peak, unit="TFLOP/s",
                        frac=achieved_tf / tensor_peak, **common)
    else:
        roofline = dict(bound="hbm", achieved=achieved_gbs, peak=hbm_peak, unit="GB/s", frac=achieved_gbs / hbm_peak,
                        **common)

    # ---- tensor mode: one extra chain in all-fp32 parity mode, reported next to the headline
    parity_leg = None
    if args.mode == "tensor" and not args.no_parity_leg and world == 1:
        pnet = bdiff.GCPNetDynamicsB200(config=dcfg, mode="parity")
        pnet.load_state_dict(sd, strict=True)
        pnet.to(dev)
        psampler = bdiff.GCDMSampler(pnet, use_cuda_graph=True)
        psampler.sample(num_nodes_dev, ctx_dev, min(T, 50))          # warm-up / graph capture
        torch.cuda.synchronize()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ev0.record()
        psampler.sample(num_nodes_dev, ctx_dev, T)
        ev1.record()
        torch.cuda.synchronize()
        psecs = ev0.elapsed_time(ev1) / 1000.0
        parity_leg = {"value": batch / psecs, "unit": "molecules/s", "ms_per_step": 1000 * psecs, "dtype": "f32",
                      "note": "same workload with BDIFF_MODE_PARITY_FP32 (every MAC an fp32 FFMA; 1e-6 from the reference)"}
        log(f"parity-mode chain done: {psecs:.2f} s")

    line = {
        "metric": METRIC, "value": value, "unit": "molecules/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1000 * secs / args.steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f32" if args.mode == "parity" else "bf16",
        "data": "synthetic",
        "config": {"workload": f"{args.config} unconditional sampling, T={T}, batch {batch} x {atoms} atoms per GPU"
                               if not dcfg.num_context else
                               f"{args.config} property-conditional sampling, T={T}, batch {batch} x {atoms} atoms per GPU",
                   "config_name": args.config, "molecules_per_gpu": batch, "atoms_per_molecule": atoms, "timesteps": T,
                   "denoiser_forwards_per_step": T + 1, "nodes_per_gpu": n_nodes, "edges_per_gpu": E,
                   "mode": args.mode,
                   "precision": ("tensor mode: bf16 GEMM operands on tcgen05 tensor cores, fp32 accumulation (TMEM) and fp32 "
                                 "state; per-forward error vs the reference <= 2e-2*max|out| (measured ~2e-3)")
                                if args.mode == "tensor" else "parity mode: all fp32 FFMA, 1e-6 from the reference",
                   "weights": "random init of the named architecture (seed 7)",
                   "l2": "flushed between timed chains (256 MiB write); inside a chain the working set is "
                         "L2-resident by design",
                   "parallelism": f"dp{world}: molecule shards, no collective in the chain, one final all_gather"},
        "e2e": {"value": e2e_value, "unit": "molecules/s", "h2d_bytes_per_step": int(num_nodes_host.numel() * 8 +
                (ctx_host.numel() * 4 if ctx_host is not None else 0)),
                "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": 1000 * secs_e2e / args.steps},
        "gpu_launches": int(launches) * world,
        "clocks": clk,
        "roofline": roofline,
    }
    if parity_leg is not None:
        line["parity_fp32"] = parity_leg
    if world == 1 and rank == 0 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_block(args)
    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def sampler_launches(sampler, net):
    return getattr(sampler, "kernel_launches", 0)


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
