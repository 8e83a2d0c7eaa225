#!/usr/bin/env python
"""tools/ncu_summary.py -- numbers for bench.py's roofline keys out of committed ncu captures.

    python tools/ncu_summary.py KEY AGENTS FULL.ncu-rep LAUNCHES.csv [--kernel step_tile_kernel]

FULL.ncu-rep : one `ncu --set full --import-source on` capture of the dominant kernel (tools/gpurun/r2_prof_tile.sh)
LAUNCHES.csv : `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum` launch list of a
               bench.py run (tools/gpurun/r2_launches.sh)
Writes / updates profiles/kernel_counts.json[KEY] = {dp_inst_per_agent, kernel_dram_bytes, step_dram_bytes, ...}.
FP64 instructions are counted per THREAD from the capture's SASS page (DADD, DMUL, DFMA, DSETP and the MUFU.*64H
seeds of sqrt / reciprocal), i.e. what the FP64 pipe executes.
"""
import csv
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def sass_counts(rep):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True,
                         text=True).stdout
    hdr, dp_thread, dp_warp, all_warp, all_thread = None, 0, 0, 0, 0
    for r in csv.reader(out.splitlines()):
        if r and r[0] == "Address":
            hdr = r
            continue
        if not hdr or len(r) != len(hdr):
            continue
        d = dict(zip(hdr, r))
        try:
            w, t = int(d["Instructions Executed"]), int(d["Thread Instructions Executed"])
        except ValueError:
            continue
        op = d["Source"].split()
        op = op[1] if op and op[0].startswith("@") else (op[0] if op else "")
        all_warp += w
        all_thread += t
        if op.startswith(("DADD", "DMUL", "DFMA", "DSETP")) or (op.startswith("MUFU") and "64H" in op):
            dp_thread += t
            dp_warp += w
    return dp_thread, dp_warp, all_warp, all_thread


def raw_metric(rep, name):
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    d = dict(zip(rows[0], rows[2]))
    return float(d[name].replace(",", ""))


def launch_list(path, kernel):
    rows = list(csv.reader(open(path)))
    hi = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
    hdr = rows[hi]
    ki, mi, vi, ii, ui = (hdr.index(k) for k in ("Kernel Name", "Metric Name", "Metric Value", "ID", "Metric Unit"))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1.0, "us": 1e3, "ms": 1e6}
    by = {}
    for r in rows[hi + 2:]:
        if len(r) <= vi:
            continue
        e = by.setdefault(int(r[ii]), {"name": r[ki].split("(")[0]})
        e[r[mi]] = float(r[vi].replace(",", "")) * scale.get(r[ui], 1.0)
    L = [by[k] for k in sorted(by)]
    begins = [i for i, x in enumerate(L) if "begin_step" in x["name"]]
    a, b = begins[-2], begins[-1]  # the last complete step
    step = [x for x in L[a:b] if "flush_l2" not in x["name"]]
    tot_ns = sum(x.get("gpu__time_duration.sum", 0.0) for x in step)
    dram = sum(x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0) for x in step)
    table = [(x["name"].split("::")[-1], x.get("gpu__time_duration.sum", 0.0) / 1e3,
              (x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0)) / 1e6) for x in step]
    k = [x for x in step if kernel in x["name"]]
    k_ns = sum(x.get("gpu__time_duration.sum", 0.0) for x in k)
    k_dram = sum(x.get("dram__bytes_read.sum", 0.0) + x.get("dram__bytes_write.sum", 0.0) for x in k)
    return tot_ns, dram, k_ns, k_dram, table


def main():
    key, agents, rep, launches = sys.argv[1], int(sys.argv[2]), sys.argv[3], sys.argv[4]
    kernel = sys.argv[sys.argv.index("--kernel") + 1] if "--kernel" in sys.argv else "step_tile_kernel"
    dp_t, dp_w, all_w, all_t = sass_counts(rep)
    tot_ns, dram, k_ns, k_dram, table = launch_list(launches, kernel)
    entry = {
        "source": f"profiles/{os.path.basename(launches)} + ncu --set full capture of {kernel} (tools/ncu_summary.py)",
        "agents": agents, "dp_inst_per_agent": dp_t / agents, "dp_warp_inst_per_32_agents": dp_w / (agents / 32),
        "warp_inst_per_32_agents": all_w / (agents / 32), "active_lanes_per_inst": all_t / max(all_w, 1),
        "kernel_dram_bytes": k_dram, "step_dram_bytes": dram, "kernel_share_of_step_ncu": k_ns / max(tot_ns, 1),
        "kernel_us_ncu": k_ns / 1e3, "step_us_ncu": tot_ns / 1e3,
        "fp64_pipe_pct": raw_metric(rep, "sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"),
        "issue_active_pct": raw_metric(rep, "smsp__issue_active.avg.pct_of_peak_sustained_active"),
    }
    path = os.path.join(ROOT, "profiles", "kernel_counts.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    data[key] = entry
    with open(path, "w") as f:
        json.dump(data, f, indent=1, sort_keys=True)
    print(json.dumps(entry, indent=1))
    print("| kernel | us | DRAM MB |\n|---|---|---|")
    for name, us, mb in table:
        print(f"| {name} | {us:.1f} | {mb:.1f} |")


if __name__ == "__main__":
    main()
