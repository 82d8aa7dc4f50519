#!/usr/bin/env python
"""Summarise ncu artefacts (read here, on the CPU box) into profiles/.
  python tools/summarise_ncu.py launches gpurun_out/launches.csv         -> per-kernel launch list
  python tools/summarise_ncu.py report gpurun_out/prof_X.ncu-rep         -> key metrics of a --set full capture
  python tools/summarise_ncu.py traffic name=gpurun_out/prof_X.ncu-rep ... -> profiles/traffic.json: DRAM bytes of one launch per
                                                                             named kernel + the commit, read by bench.py
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_warps",
    "gpu__time_duration.sum", "sm__cycles_elapsed.max", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "dram__throughput.avg.pct_of_peak_sustained_elapsed", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "lts__t_bytes.sum", "lts__throughput.avg.pct_of_peak_sustained_elapsed", "l1tex__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_short_scoreboard_per_warp_active.pct",
    "smsp__warp_issue_stalled_math_pipe_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_not_selected_per_warp_active.pct",
    "smsp__warp_issue_stalled_barrier_per_warp_active.pct", "smsp__warp_issue_stalled_wait_per_warp_active.pct",
    "smsp__warp_issue_stalled_lg_throttle_per_warp_active.pct", "smsp__warp_issue_stalled_mio_throttle_per_warp_active.pct",
    "smsp__warp_issue_stalled_dispatch_stall_per_warp_active.pct", "smsp__warp_issue_stalled_branch_resolving_per_warp_active.pct",
    "smsp__warp_issue_stalled_no_instruction_per_warp_active.pct", "smsp__warp_issue_stalled_selected_per_warp_active.pct",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    by_grid = collections.OrderedDict()
    for row in csv.DictReader(lines):
        if row.get("Metric Name") != "gpu__time_duration.sum":
            continue
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v = v / 1e3 if u in ("nsecond", "ns") else (v * 1e3 if u in ("msecond", "ms") else v)
        agg.setdefault(name, []).append(v)
        by_grid.setdefault((name, row.get("Grid Size", "")), []).append(v)
    tot = sum(sum(v) for v in agg.values())
    print(f"{'kernel':72s} {'n':>5s} {'total_us':>12s} {'avg_us':>10s} {'share':>6s}")
    for k, v in sorted(agg.items(), key=lambda kv: -sum(kv[1])):
        print(f"{k[:72]:72s} {len(v):5d} {sum(v):12.1f} {sum(v) / len(v):10.1f} {sum(v) / tot:6.3f}")
    # kernels launched with several grid sizes (e.g. the full-stripe launches of the timed steps vs the chunked launches
    # of emo_mosaic): one line per grid
    multi = {k for k, _ in by_grid if sum(1 for kk, _ in by_grid if kk == k) > 1}
    if multi:
        print()
        print(f"{'kernel, by grid size':60s} {'grid':>18s} {'n':>5s} {'avg_us':>10s}")
        for (k, g), v in by_grid.items():
            if k in multi:
                print(f"{k[:60]:60s} {g:>18s} {len(v):5d} {sum(v) / len(v):10.1f}")


def report(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    for vals in rows[2:]:
        for w in KEYS:
            if w in hdr:
                i = hdr.index(w)
                print(f"{w:72s} {vals[i]:>24s} {units[i]}")
        print()


UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
TIME = {"nsecond": 1e-3, "ns": 1e-3, "usecond": 1.0, "us": 1.0, "msecond": 1e3, "ms": 1e3, "second": 1e6, "s": 1e6}


def traffic(pairs):
    """name=report pairs -> profiles/traffic.json {commit, kernels: {name: {dram_bytes, dram_read, dram_write, ncu_us, kernel, report}}}"""
    import json
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    commit = subprocess.run(["git", "-C", root, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = {"commit": commit, "source": "ncu --set full --clock-control none, one launch each (tools/profile_r02.sh)", "kernels": {}}
    for pair in pairs:
        name, path = pair.split("=", 1)
        txt = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
        rows = list(csv.reader(txt.splitlines()))
        if len(rows) < 3:
            print(f"{name}: no data in {path}")
            continue
        hdr, units, vals = rows[0], rows[1], rows[2]

        def get(metric, table):
            i = hdr.index(metric)
            return float(vals[i].replace(",", "")) * table[units[i]]

        rd, wr = get("dram__bytes_read.sum", UNIT), get("dram__bytes_write.sum", UNIT)
        out["kernels"][name] = {"dram_bytes": int(rd + wr), "dram_read": int(rd), "dram_write": int(wr),
                                "ncu_us": get("gpu__time_duration.sum", TIME), "kernel": vals[hdr.index("Kernel Name")],
                                "report": os.path.basename(path)}
        print(name, out["kernels"][name])
    json.dump(out, open(os.path.join(root, "profiles", "traffic.json"), "w"), indent=1)


if __name__ == "__main__":
    if sys.argv[1] == "traffic":
        traffic(sys.argv[2:])
    else:
        {"launches": launches, "report": report}[sys.argv[1]](sys.argv[2])
