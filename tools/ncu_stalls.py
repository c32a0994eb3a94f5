#!/usr/bin/env python
"""Aggregate the per-instruction stall samples of an `ncu --page source --csv` dump.
usage: ncu -i rep.ncu-rep --page source --csv > src.csv; python tools/ncu_stalls.py src.csv [kernel-index]"""
import csv
import re
import sys
from collections import Counter


def main(path, which=0):
    rows = list(csv.reader(open(path)))
    # the dump holds one table per kernel: a "Kernel Name" row, a header row, then instructions
    tables, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "rows": []}
            tables.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and r:
            cur["rows"].append(r)
    t = tables[which]
    hdr = t["hdr"]
    stall_cols = [i for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    isamp, iexec, isrc = hdr.index("# Samples"), hdr.index("Instructions Executed"), hdr.index("Source")
    tot = Counter()
    opc = Counter()
    opc_exec = Counter()
    total_samples = 0
    total_exec = 0
    for r in t["rows"]:
        try:
            n = int(r[isamp])
        except ValueError:
            continue
        total_samples += n
        ex = int(r[iexec])
        total_exec += ex
        op = r[isrc].split()[0] if r[isrc].split() else "?"
        if op.startswith("@"):
            op = r[isrc].split()[1]
        op = re.sub(r"\..*", "", op)
        opc[op] += n
        opc_exec[op] += ex
        for i in stall_cols:
            tot[hdr[i]] += int(r[i] or 0)
    print(t["name"][:100])
    print("instructions:", len(t["rows"]), " samples:", total_samples, " warp-instr executed:", total_exec)
    print("stall reasons (share of samples):")
    for k, v in tot.most_common(10):
        print(f"   {k:28s} {100.0 * v / max(1, total_samples):5.1f}%")
    print("by opcode: samples%  exec%")
    for k, v in opc.most_common(14):
        print(f"   {k:10s} {100.0 * v / max(1, total_samples):5.1f}%  {100.0 * opc_exec[k] / max(1, total_exec):5.1f}%")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
