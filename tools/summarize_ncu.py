#!/usr/bin/env python
"""Turn gpurun_out/ ncu artefacts into the tracked summaries under profiles/.

  python tools/summarize_ncu.py <tag> [--launches gpurun_out/launches.csv] [--rep gpurun_out/prof_scan.ncu-rep]

Writes profiles/<tag>_launches.md (per-kernel count / avg time / share of the step),
profiles/<tag>_scan_full.csv (key metrics of the dominant kernel from the --set full
capture) and refreshes profiles/scan_traffic.json (dram bytes per launch, read by bench.py).
"""
import argparse
import collections
import csv
import json
import os
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__bytes_read.sum.per_second", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "launch__waves_per_multiprocessor", "smsp__inst_executed.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__cycles_elapsed.max",
        "smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio",
        "smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio"]


def launches(path, tag, note):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        agg.setdefault(row["Kernel Name"], []).append(float(row["Metric Value"].replace(",", "")))
    ours = {k: v for k, v in agg.items()
            if any(n in k for n in ("scan_topk_kernel", "epilogue_kernel", "append_kernel", "reset_kernel", "log_update_kernel"))}
    per_step = sum(sum(v) / len(v) for v in ours.values())
    out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)", "", note, "",
           "Per-launch times under ncu are cold-cache and serialised: compare SHARES, not absolutes.", "",
           "| kernel | launches | avg us | share of one step (ours) |", "|---|---:|---:|---:|"]
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        avg = sum(v) / len(v) / 1e3
        share = f"{(sum(v) / len(v)) / per_step * 100:.1f}%" if k in ours else "(setup, torch)"
        out.append(f"| `{k[:90]}` | {len(v)} | {avg:.2f} | {share} |")
    open(os.path.join(ROOT, "profiles", f"{tag}_launches.md"), "w").write("\n".join(out) + "\n")
    print("\n".join(out))


def full(rep, tag):
    raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(raw.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    out_rows = [["metric", "unit"] + [f"launch{i}" for i in range(len(data))]]
    vals = {}
    for key in KEYS:
        if key in hdr:
            i = hdr.index(key)
            out_rows.append([key, units[i]] + [r[i] for r in data])
            vals[key] = [r[i] for r in data]
    with open(os.path.join(ROOT, "profiles", f"{tag}_scan_full.csv"), "w", newline="") as f:
        csv.writer(f).writerows(out_rows)

    def mb(key):
        i = hdr.index(key)
        scale = {"Mbyte": 1e6, "Gbyte": 1e9, "Kbyte": 1e3, "byte": 1.0}[units[i]]
        xs = [float(r[i].replace(",", "")) * scale for r in data]
        return sum(xs) / len(xs)
    traffic = mb("dram__bytes_read.sum") + mb("dram__bytes_write.sum")
    json.dump({"dram_bytes_per_launch": traffic, "source": f"profiles/{tag}_scan_full.csv (ncu --set full, mean of {len(data)} launches)",
               "kernel": vals["Kernel Name"][0]}, open(os.path.join(ROOT, "profiles", "scan_traffic.json"), "w"), indent=1)
    for r in out_rows:
        print(r)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("tag")
    ap.add_argument("--launches", default=os.path.join(ROOT, "gpurun_out", "launches.csv"))
    ap.add_argument("--rep", default=os.path.join(ROOT, "gpurun_out", "prof_scan.ncu-rep"))
    ap.add_argument("--note", default="command: python bench.py --steps 5 --warmup 3 --no-cpu-baseline")
    a = ap.parse_args()
    if os.path.exists(a.launches):
        launches(a.launches, a.tag, a.note)
    if os.path.exists(a.rep):
        full(a.rep, a.tag)
