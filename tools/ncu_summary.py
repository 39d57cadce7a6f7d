"""Summarise ncu evidence into markdown for profiles/ (run here, no GPU needed).

  python tools/ncu_summary.py launches <launches.csv> [title]      per-kernel totals / shares of a launch list
  python tools/ncu_summary.py report <file.ncu-rep> [title]        key metrics of every captured launch
  python tools/ncu_summary.py traffic <launches.csv> <workload>    DRAM bytes per kernel family (-> profiles/traffic.json)
"""
import csv
import io
import re
import subprocess
import sys
from collections import OrderedDict

KEYS = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "regs/thread"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy %"),
    ("smsp__thread_inst_executed_per_inst_executed.ratio", "threads active per instruction (of 32)"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue slot utilisation %"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput %"),
    ("l1tex__throughput.avg.pct_of_peak_sustained_active", "L1/TEX throughput %"),
    ("lts__throughput.avg.pct_of_peak_sustained_elapsed", "L2 throughput %"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput %"),
    ("dram__bytes_read.sum", "DRAM read"),
    ("dram__bytes_write.sum", "DRAM write"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__t_sector_hit_rate.pct", "L1 hit %"),
    ("lts__t_sector_hit_rate.pct", "L2 hit %"),
    ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_scoreboard"),
    ("smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio", "stall no_instruction"),
    ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait"),
    ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_scoreboard"),
    ("smsp__average_warps_issue_stalled_branch_resolving_per_issue_active.ratio", "stall branch_resolving"),
    ("smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio", "stall lg_throttle"),
    ("smsp__average_warps_issue_stalled_math_pipe_throttle_per_issue_active.ratio", "stall math_pipe_throttle"),
    ("smsp__average_warps_issue_stalled_not_selected_per_issue_active.ratio", "stall not_selected"),
    ("smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio", "stall mio_throttle"),
]


def short(name):
    name = re.sub(r"^void\s+", "", name)
    name = re.sub(r"slrgpu::", "", name)
    return re.sub(r"\(.*$", "", name)


BYTE_UNITS = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}


def read_launches(path):
    """{kernel: [launches, total us, dram bytes]} from an `ncu --csv` launch list (one row per launch and metric)."""
    rows = [r for r in csv.reader(open(path)) if r]
    hdr_i = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    hdr = rows[hdr_i]
    kn, mn, mv, mu = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    tot = OrderedDict()
    for r in rows[hdr_i + 1:]:
        if len(r) <= mv:
            continue
        v = float(r[mv].replace(",", ""))
        unit = r[mu]
        t = tot.setdefault(short(r[kn]), [0, 0.0, 0.0])
        if r[mn] == "gpu__time_duration.sum":
            t[0] += 1
            t[1] += v / 1e3 if unit in ("ns", "nsecond") else v if unit in ("us", "usecond") else v * 1e3 if unit in ("ms", "msecond") else v
        elif r[mn].startswith("dram__bytes"):
            t[2] += v * BYTE_UNITS.get(unit, 1.0)
    return tot


def launches(path, title):
    tot = read_launches(path)
    total = sum(t[1] for t in tot.values())
    print(f"## {title}\n\n`{path}`: {sum(t[0] for t in tot.values())} launches, {total / 1e3:.2f} ms of kernel time "
          f"(serialised, cold-cache: compare shares)\n")
    print("| kernel | launches | total ms | share | mean us | DRAM read+write MB (all launches) |\n|---|---|---|---|---|---|")
    for k, t in sorted(tot.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {t[0]} | {t[1] / 1e3:.3f} | {100 * t[1] / total:.1f} % | {t[1] / max(t[0], 1):.1f} | {t[2] / 1e6:.1f} |")


def traffic(path, workload):
    """Per kernel family DRAM bytes of the listed launches (one frame): the `traffic` of bench.py's roofline object."""
    fam = {}
    for k, t in read_launches(path).items():
        name = re.sub(r"<.*$", "", k)
        if name.endswith("Kernel") and name[:-6] in ("extend", "shadow", "surface", "material", "raygen", "tail"):
            f = fam.setdefault(name, {"dram_bytes_per_frame": 0.0, "launches": 0, "kernel_ms_per_frame_under_ncu": 0.0})
            f["dram_bytes_per_frame"] += t[2]
            f["launches"] += t[0]
            f["kernel_ms_per_frame_under_ncu"] += t[1] / 1e3
    import json
    print(json.dumps({workload: fam}, indent=1))


def report(path, title):
    if path.endswith(".csv.gz"):      # the raw page exported on the GPU box
        import gzip
        out = gzip.open(path, "rt").read()
    else:
        out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    print(f"## {title}\n\n`{path}` (`ncu --set full --clock-control none`), one column per captured launch\n")
    names = [short(r[hdr.index("Kernel Name")]) for r in rows[2:]]
    print("| metric | " + " | ".join(f"`{n}`" for n in names) + " |")
    print("|---|" + "---|" * len(names))
    for key, label in KEYS:
        if key not in hdr:
            continue
        i = hdr.index(key)
        vals = []
        for r in rows[2:]:
            try:
                vals.append(f"{float(r[i].replace(',', '')):.4g}")
            except ValueError:
                vals.append(r[i])
        print(f"| {label} [{units[i]}] | " + " | ".join(vals) + " |")


if __name__ == "__main__":
    mode, path = sys.argv[1], sys.argv[2]
    title = sys.argv[3] if len(sys.argv) > 3 else path
    {"launches": launches, "report": report, "traffic": traffic}[mode](path, title)
