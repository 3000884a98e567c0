#!/usr/bin/env python
"""Where a warp's time goes inside k_extend_sym, from the source page of an `ncu --set full --import-source on`
capture: per captured launch the stall-reason shares of the warp samples, the samples by SASS opcode, and the
instructions that collect the long-scoreboard (global load / mbarrier wait) samples.

  tools/ncu_source_summary.py <file.ncu-rep>  > profiles/rNN_x_ncu_source_stalls.txt
"""
import collections
import csv
import io
import subprocess
import sys


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    blocks = txt.split('"Kernel Name"')
    seen = set()
    print(f"# {rep}: warp-state samples per SASS instruction (ncu source page), one block per captured launch")
    for b in blocks[1:]:
        lines = b.split("\n")
        name = lines[0].strip().strip(",").strip('"')
        rdr = csv.reader(io.StringIO("\n".join(lines[1:])))
        hdr = next(rdr)
        idx = {h: i for i, h in enumerate(hdr)}
        rows = [r for r in rdr if len(r) == len(hdr)]
        tot = sum(int(r[idx["# Samples"]]) for r in rows)
        ex = sum(int(r[idx["Instructions Executed"]]) for r in rows)
        key = (len(rows), tot, ex)
        if key in seen or tot == 0:
            continue
        seen.add(key)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = collections.Counter()
        for r in rows:
            for s in stalls:
                agg[s] += int(r[idx[s]])
        print(f"\n## {name[:90]}: {len(rows)} SASS instructions, {ex} warp instructions executed, {tot} samples")
        print("stall shares: " + ", ".join(f"{k[6:]} {v / tot:.1%}" for k, v in agg.most_common(9)))
        byop, cnt = collections.Counter(), collections.Counter()
        for r in rows:
            parts = r[idx["Source"]].split()
            op = parts[1] if parts and parts[0].startswith("@") and len(parts) > 1 else (parts[0] if parts else "?")
            byop[op] += int(r[idx["# Samples"]])
            cnt[op] += int(r[idx["Instructions Executed"]])
        print("samples by opcode: " + ", ".join(f"{k} {v / tot:.1%} ({cnt[k] / ex:.1%} of instructions)" for k, v in byop.most_common(8)))
        top = sorted(rows, key=lambda r: -int(r[idx["stall_long_sb"]]))[:5]
        print("long-scoreboard samples collect at: " + "; ".join(
            f"{r[idx['Source']].strip()[:44]} [{int(r[idx['stall_long_sb']]) / tot:.1%}]" for r in top if int(r[idx["stall_long_sb"]])))


if __name__ == "__main__":
    main()
