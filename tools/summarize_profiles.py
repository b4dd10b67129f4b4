#!/usr/bin/env python
"""Turns the ncu outputs brought back in gpurun_out/ into the tracked summaries under profiles/.
  launches CSV (ncu --metrics gpu__time_duration.sum)  -> profiles/<tag>_launches.md
  .ncu-rep (ncu --set full, one kernel)                -> profiles/<tag>_ncu_<kernel>.md (+ .csv of the metrics used)
Usage: python tools/summarize_profiles.py <tag> <launches.csv> <kernel-name> <file.ncu-rep>"""
import collections
import csv
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEEP = ["gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__shared_mem_per_block_dynamic",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_bytes.sum",
        "lts__throughput.avg.pct_of_peak_sustained_elapsed", "lts__t_sector_hit_rate.pct", "l1tex__t_sector_hit_rate.pct",
        "smsp__warps_eligible.avg.per_cycle_active", "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
        "smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_dispatch_stall_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "smsp__thread_inst_executed_per_inst_executed.ratio"]


def launches(tag, fn):
    rows = [r for r in csv.reader(open(fn)) if len(r) > 10]
    hdr = rows[0]
    ix = {h: i for i, h in enumerate(hdr)}
    agg = collections.OrderedDict()
    for r in rows[1:]:
        try:
            v = float(r[ix["Metric Value"]].replace(",", ""))
        except ValueError:
            continue
        k = (r[ix["Kernel Name"]].split("(")[0].replace("void ", ""), r[ix["Block Size"]], r[ix["Grid Size"]])
        a = agg.setdefault(k[0], [0, 0.0, set()])
        a[0] += 1; a[1] += v; a[2].add(f"{k[2]}x{k[1]}")
    tot = sum(a[1] for a in agg.values())
    out = [f"# {tag}: kernel launch list (ncu --metrics gpu__time_duration.sum --clock-control none)", "",
           f"source: `{os.path.basename(fn)}`; per-launch times are cold-cache and serialised: compare SHARES, not absolutes", "",
           "| kernel | launches | total ms | share | grid x block |", "|---|---:|---:|---:|---|"]
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        out.append(f"| `{k}` | {a[0]} | {a[1] / 1e6:.3f} | {100 * a[1] / tot:.1f}% | {', '.join(sorted(a[2]))[:60]} |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def ncu(tag, kernel, rep):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units = rows[0], rows[1]
    ki = hdr.index("Kernel Name")
    vals = next((r for r in rows[2:] if kernel in r[ki]), rows[2])     # first profiled launch of that kernel
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    if kernel.startswith("direct_fp32"):
        # per-launch DRAM traffic of the dominant kernel: bench.py's roofline.traffic reads this file
        import json
        def _bytes(key):
            v, u = d[key]
            return float(v.replace(",", "")) * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        json.dump({"direct_fp32_kernel": {"dram_bytes_read": _bytes("dram__bytes_read.sum"),
                                          "dram_bytes_write": _bytes("dram__bytes_write.sum"),
                                          "source": f"profiles/{tag}_ncu_{kernel}.md (ncu --set full, one launch of 800k poses)"}},
                  open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    out = [f"# {tag}: ncu --set full --clock-control none, kernel `{kernel}`", "", f"source: `{os.path.basename(rep)}` (first profiled launch)", "",
           "| metric | value | unit |", "|---|---:|---|"]
    keep_rows = []
    for k in KEEP:
        if k in d:
            out.append(f"| {k} | {d[k][0]} | {d[k][1]} |")
            keep_rows.append((k, d[k][0], d[k][1]))
    # per-phase split of the warp-stall samples (tools/ncu_phases.py: segments by execution count)
    ph = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "ncu_phases.py"), rep, kernel, "12"], capture_output=True, text=True).stdout
    if ph.strip():
        out += ["", "## where the warps spend their time (source page, stall samples per code segment)", "", "```", ph.rstrip(), "```"]
    open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{kernel}.md"), "w").write("\n".join(out) + "\n")
    with open(os.path.join(ROOT, "profiles", f"{tag}_ncu_{kernel}.csv"), "w") as f:
        w = csv.writer(f)
        w.writerow(["metric", "value", "unit"])
        w.writerows(keep_rows)
    print("\n".join(out))


if __name__ == "__main__":
    tag = sys.argv[1]
    if sys.argv[2] != "-":
        launches(tag, sys.argv[2])
    if len(sys.argv) > 4:
        for kern in sys.argv[3].split(","):
            ncu(tag, kern, sys.argv[4])
