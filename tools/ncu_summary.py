#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py shares <launches.csv> <out.txt> "<command line that was profiled>"
  python tools/ncu_summary.py full   <report.ncu-rep> <out.txt> "<command line>"   (needs ncu on PATH; no GPU needed)
"""
import collections
import csv
import subprocess
import sys

KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__shared_mem_per_block_dynamic", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers",
        "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__inst_executed.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__pipe_tensor_subpipe_hmma_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_sleeping_per_issue_active.ratio"]


def shares(path, out, cmd):
    rows = list(csv.reader(open(path)))
    hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    H = rows[hdr]
    ki, vi = H.index("Kernel Name"), H.index("Metric Value")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[hdr + 1:]:
        if len(r) <= vi:
            continue
        name = r[ki].split("(")[0].replace("bdiff::", "").replace("void ", "")
        agg[name][0] += 1
        agg[name][1] += float(r[vi].replace(",", ""))
    tot = sum(v[1] for v in agg.values())
    lines = [f"# {cmd}", "# ncu --metrics gpu__time_duration.sum --clock-control none: per-launch times are cold-cache and "
             "serialised -> compare SHARES, not absolutes",
             f"{'kernel':36s} {'launches':>8s} {'total_us':>10s} {'avg_us':>9s} {'share':>7s}"]
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append(f"{k:36s} {v[0]:8d} {v[1] / 1e3:10.1f} {v[1] / 1e3 / v[0]:9.1f} {100 * v[1] / tot:6.1f}%")
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


def full(rep, out, cmd):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    H, units = rows[0], rows[1]
    lines = [f"# {cmd}", "# ncu --set full --clock-control none --import-source on (one launch); values as reported by ncu"]
    for r in rows[2:3]:
        lines.append(f"kernel: {r[H.index('Kernel Name')]}")
        for k in KEEP:
            if k in H:
                lines.append(f"{k:95s} {r[H.index(k)]:>20s} {units[H.index(k)]}")
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    srows = list(csv.reader(src.splitlines()))
    if len(srows) > 2:
        SH = srows[1]
        si = SH.index("# Samples")
        cols = [i for i, h in enumerate(SH) if h.startswith("stall_") and "Not Issued" not in h]
        tot = collections.Counter()
        n = 0
        for r in srows[2:]:
            try:
                n += int(r[si])
            except (ValueError, IndexError):
                continue
            for i in cols:
                try:
                    tot[SH[i]] += int(r[i])
                except ValueError:
                    pass
        lines.append(f"warp-state samples: {n}; by stall reason: " + ", ".join(f"{k}={v}" for k, v in tot.most_common(10)))
    open(out, "w").write("\n".join(lines) + "\n")
    print("\n".join(lines))


if __name__ == "__main__":
    {"shares": shares, "full": full}[sys.argv[1]](*sys.argv[2:5])
