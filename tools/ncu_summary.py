#!/usr/bin/env python
"""Summarise ncu outputs into small text files for profiles/.
  tools/ncu_summary.py launches <launches.csv>          -> per-kernel totals and shares
  tools/ncu_summary.py full <file.ncu-rep> [regex]      -> key metrics per captured launch
"""
import collections
import csv
import re
import subprocess
import sys

KEYS = [
    "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread",
    "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
    "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active",
    "smsp__issue_active.avg.pct_of_peak_sustained_active",
    "sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
    "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active",
    "smsp__inst_executed.sum", "smsp__cycles_active.avg",
    "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
    "lts__t_sector_hit_rate.pct",
]


def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    tot, cnt = collections.defaultdict(float), collections.Counter()
    for row in csv.DictReader(lines):
        if row.get("Metric Name", "gpu__time_duration.sum") != "gpu__time_duration.sum":
            continue
        v = float(row["Metric Value"].replace(",", ""))
        u = row["Metric Unit"]
        v *= {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0, "s": 1e3, "second": 1e3}.get(u, 1e-6)
        name = re.sub(r"\(.*", "", row["Kernel Name"])
        name = re.sub(r".*k_map<.*?(\w+)\(.*", r"k_map<\1>", row["Kernel Name"]) if "k_map" in row["Kernel Name"] else name
        tot[name] += v
        cnt[name] += 1
    s = sum(tot.values())
    print(f"# {path}: {sum(cnt.values())} launches, {s:.2f} ms total device time (cold-cache, serialised: compare SHARES)")
    for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
        print(f"{v:10.3f} ms {cnt[k]:6d} launches {100 * v / s:5.1f}%  {k}")


def full(path, pattern=None):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}
    print(f"# {path}: {len(data)} captured launches")
    print("kernel:", [re.sub(r"\(.*", "", r[idx["Kernel Name"]]) for r in data])
    for k in KEYS:
        if k in idx:
            print(f"{k:70s} [{units[idx[k]]}] " + "  ".join(r[idx[k]] for r in data))
    for h in hdr:
        if (h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio")) or "fmaheavy_cycles_active.avg" in h or "fmalite_cycles_active.avg" in h or (pattern and re.search(pattern, h)):
            print(f"{h:70s} [{units[idx[h]]}] " + "  ".join(r[idx[h]] for r in data))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2])
    else:
        full(sys.argv[2], sys.argv[3] if len(sys.argv) > 3 else None)
