#!/usr/bin/env python
"""profiles/rNN_traffic.json from an ncu launch list of ONE operation
(ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv ... tools/one_op.py):
per kernel family: launches, summed duration and DRAM bytes.  bench.py's roofline.traffic reads it.

  tools/traffic_from_launches.py <launches.csv> <out.json> "<source note>"
"""
import collections
import csv
import json
import re
import sys


def main():
    path, out, note = sys.argv[1], sys.argv[2], sys.argv[3]
    lines = [l for l in open(path) if not l.startswith("==")]
    per = collections.defaultdict(lambda: collections.defaultdict(float))
    ids = collections.defaultdict(set)
    scale_t = {"ns": 1e-6, "nsecond": 1e-6, "us": 1e-3, "usecond": 1e-3, "ms": 1.0, "msecond": 1.0}
    scale_b = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0}
    for row in csv.DictReader(lines):
        name = row["Kernel Name"]
        fam = ("k_extend_sym" if "k_extend_sym" in name else "k_enter_combine" if "k_enter_combine" in name
               else re.sub(r"\(.*", "", name).split("::")[-1])
        v = float(row["Metric Value"].replace(",", ""))
        m, u = row["Metric Name"], row["Metric Unit"]
        ids[fam].add(row["ID"])
        if m == "gpu__time_duration.sum":
            per[fam]["ncu_time_ms"] += v * scale_t.get(u, 1e-6)
        elif m == "dram__bytes_read.sum":
            per[fam]["dram_read_gb"] += v * scale_b.get(u, 1e-9)
        elif m == "dram__bytes_write.sum":
            per[fam]["dram_write_gb"] += v * scale_b.get(u, 1e-9)
    res = {}
    for fam, d in per.items():
        res[fam] = {"launches_per_step": len(ids[fam]), "ncu_time_ms_per_step": round(d["ncu_time_ms"], 3),
                    "dram_read_gb_per_step": round(d["dram_read_gb"], 3), "dram_write_gb_per_step": round(d["dram_write_gb"], 3),
                    "dram_total_gb_per_step": round(d["dram_read_gb"] + d["dram_write_gb"], 3)}
    res["source"] = note
    json.dump(res, open(out, "w"), indent=1)
    print(json.dumps(res)[:600])


if __name__ == "__main__":
    main()
