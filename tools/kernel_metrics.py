#!/usr/bin/env python
"""profiles/kernel_metrics.json from `ncu --set full` reports: per workload and kernel family the numbers bench.py's roofline
object quotes when a kernel is issue bound -- lanes active per instruction, issue-slot utilisation, L1 hit rate, occupancy.
    python tools/kernel_metrics.py <workload> <file.ncu-rep> <source note> [<workload> <file> <note> ...]
The first captured launch of every kernel family in a report is used (the capture scripts take the first, full-pool wave)."""
import csv
import io
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = {"lanes_active": "smsp__thread_inst_executed_per_inst_executed.ratio",
        "issue_slot_utilisation_pct": "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "l1_hit_pct": "l1tex__t_sector_hit_rate.pct",
        "occupancy_pct": "sm__warps_active.avg.pct_of_peak_sustained_active",
        "dram_throughput_pct": "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "duration_ms": "gpu__time_duration.sum"}


def read(rep):
    if rep.endswith(".csv.gz"):       # the raw page exported on the GPU box (reports larger than a call may bring back)
        import gzip
        out = gzip.open(rep, "rt").read()
    else:
        out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    res = {}
    for r in rows[2:]:
        name = re.sub(r"^void\s+", "", r[hdr.index("Kernel Name")])
        fam = re.sub(r"<.*$", "", re.sub(r"slrgpu::", "", re.sub(r"\(.*$", "", name)))
        if fam in res:
            continue
        m = {}
        for k, col in KEYS.items():
            if col in hdr:
                v = float(r[hdr.index(col)].replace(",", ""))
                if k == "duration_ms":
                    u = units[hdr.index(col)]
                    v = v / 1e6 if u in ("ns", "nsecond") else v / 1e3 if u in ("us", "usecond") else v
                m[k] = round(v, 3)
        m["kernel"] = re.sub(r"slrgpu::", "", re.sub(r"\(.*$", "", name))
        res[fam] = m
    return res


def main():
    path = os.path.join(ROOT, "profiles", "kernel_metrics.json")
    data = json.load(open(path)) if os.path.exists(path) else {}
    a = sys.argv[1:]
    for i in range(0, len(a), 3):
        workload, rep, note = a[i], a[i + 1], a[i + 2]
        fams = read(rep)
        for f in fams.values():
            f["source"] = note
        data.setdefault(workload, {}).update(fams)
        print(workload, {k: (v["lanes_active"], v["issue_slot_utilisation_pct"]) for k, v in fams.items()})
    json.dump(data, open(path, "w"), indent=1)


if __name__ == "__main__":
    main()
