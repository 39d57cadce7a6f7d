"""Hot source lines of one captured kernel: joins the SASS page of an ncu report (warp-stall samples per
instruction) with nvdisasm's line table of the same kernel in the built object (needs -lineinfo).
  python tools/ncu_hot_lines.py <file.ncu-rep> <kernel name substring, demangled>[#k-th launch] <object.o> <mangled-name substring> [top N]
"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os
from collections import Counter


def main():
    rep, want, obj, sym = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4]
    top = int(sys.argv[5]) if len(sys.argv) > 5 else 40
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    starts = [i for i, r in enumerate(rows) if r and r[0] == "Kernel Name"] + [len(rows)]
    want, _, nth = want.partition("#")              # "kernelName#k": the k-th captured launch of that kernel (0-based)
    matches = [k for k, i in enumerate(starts[:-1]) if want in rows[i][1]]
    idx = matches[2 * int(nth or 0)]                # every launch has two sections (SASS, then the same with source markers)
    rows = rows[starts[idx]:starts[idx + 1]]           # the first captured launch of that kernel
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    cols = {n: i for i, n in enumerate(rows[hdr])}
    base = None
    samples = {}
    stall_cols = [n for n in cols if n.startswith("stall_") or n.startswith("Stall")]
    for r in rows[hdr + 1:]:
        if len(r) <= cols["# Samples"]:
            continue
        a = int(r[cols["Address"]], 16)
        base = a if base is None else base
        samples[a - base] = (int(r[cols["# Samples"]] or 0), r[cols["Source"]].strip(), int(r[cols["Instructions Executed"]] or 0))
    print(rows[0][1][:120] if rows and len(rows[0]) > 1 else "")
    d = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=d, capture_output=True)
    cubin = [f for f in os.listdir(d) if f.endswith(".cubin")][0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(d, cubin)], capture_output=True, text=True).stdout.splitlines()
    start = next(i for i, l in enumerate(dis) if l.startswith("//---") and sym in l)
    line = ("?", 0)
    by_line, by_file = Counter(), Counter()
    total = 0
    inst_by_line = Counter()
    for l in dis[start + 1:]:
        if l.startswith("//---"):
            break
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/", l)
        if m:
            off = int(m.group(1), 16)
            n, _, ex = samples.get(off, (0, "", 0))
            by_line[line] += n
            by_file[line[0]] += n
            inst_by_line[line] += ex
            total += n
    print(f"{total} samples")
    inst_by_file = Counter()
    for (f, ln), n in inst_by_line.items():
        inst_by_file[f] += n
    tot_inst = sum(inst_by_file.values())
    print(f"{tot_inst} warp instructions executed")
    for f, n in by_file.most_common(12):
        print(f"  {f:28s} {100 * n / max(total, 1):5.1f} % of samples  {100 * inst_by_file[f] / max(tot_inst, 1):5.1f} % of instructions")
    src_cache = {}
    root = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "slr_b200", "csrc")
    for (f, ln), n in by_line.most_common(top):
        if f not in src_cache:
            try:
                src_cache[f] = open(os.path.join(root, f)).read().splitlines()
            except OSError:
                src_cache[f] = []
        text = src_cache[f][ln - 1].strip()[:110] if 0 < ln <= len(src_cache[f]) else ""
        print(f"{100 * n / max(total, 1):5.1f} %  {inst_by_line[(f, ln)]:9d} inst  {f}:{ln}  {text}")


if __name__ == "__main__":
    main()
