"""SASS instruction mix of every kernel in the built library (no GPU needed):
    python tools/sass_mix.py r02 > profiles/r02_sass_mix.md    # also rewrites profiles/sass/r02_<kernel>.sass for the ray kernels
"""
import os
import re
import subprocess
import sys
from collections import Counter

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "slr_b200", "lib", "libslrgpu.so")
LISTED = ("extendKernel<false, false, false>", "extendKernel<true, false, false>", "shadowKernel<false, 16, false, false>",
          "intersectBatchKernel<false, false, false>", "surfaceKernel<16>", "raygenKernel<16>")
TAG = sys.argv[1] if len(sys.argv) > 1 else "r02"


def demangle(name):
    return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()


def short(d):
    d = re.sub(r"\(.*$", "", d).replace("void ", "").replace("slrgpu::", "")
    return d.replace("(bool)0", "false").replace("(bool)1", "true").replace("(int)", "")


def main():
    text = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    funcs = {}
    cur = None
    for line in text.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = short(demangle(m.group(1)))
            funcs[cur] = []
            continue
        if cur is not None and re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+\S", line):
            funcs[cur].append(line)
    print(f"# {TAG} -- SASS instruction mix of the kernels (final pipeline of the round; `cuobjdump -sass slr_b200/lib/libslrgpu.so`, sm_100a; `tools/sass_mix.py`)\n")
    print(f"Full listings of the ray kernels: `profiles/sass/{TAG}_*.sass`. The shade kernels and the tail kernel are 10-60 k instructions each "
          "including their out-of-line device functions (16-wavelength code, all texture kinds reachable), so only their mix is listed.\n")
    print("| kernel | instructions | top opcodes |\n|---|---|---|")
    tc = 0
    for name in sorted(funcs):
        if "Kernel" not in name or "<3" in name or ", 3>" in name and "shadow" in name:
            continue
        ops = Counter()
        for l in funcs[name]:
            body = re.sub(r"/\*[0-9a-f]{4,}\*/", "", l, count=1).strip()
            tok = body.split()
            op = tok[1] if tok and tok[0].startswith("@") and len(tok) > 1 else (tok[0] if tok else "")
            ops[op.split(".")[0].rstrip(";")] += 1
        tc += sum(n for o, n in ops.items() if o.startswith(("HMMA", "UTC", "UTMA", "QMMA")))
        top = ", ".join(f"{o} {n}" for o, n in ops.most_common(8))
        print(f"| `{name}` | {len(funcs[name])} | {top} |")
    print(f"\nTensor-core (`UTC*MMA`, `HMMA`) or TMA (`UTMALDG`) instructions in the library: {tc} -- by design: nothing on this path is a dense "
          "contraction, and all global traffic is 16-byte `LDG.E.128` / `STG.E.128` through the SoA queues plus `RED.E.ADD.F32x4` for the sensor splat.")
    out = os.path.join(ROOT, "profiles", "sass")
    os.makedirs(out, exist_ok=True)
    for name in LISTED:
        if name in funcs:
            stem = re.sub(r"<.*", "", name) + ("_instanced" if name.startswith("extendKernel<true") else "")
            with open(os.path.join(out, f"{TAG}_{stem}.sass"), "w") as f:
                f.write(f"// {name}, sm_100a, final pipeline of {TAG}\n" + "\n".join(funcs[name]) + "\n")


if __name__ == "__main__":
    main()
