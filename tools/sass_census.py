#!/usr/bin/env python3
"""Per-kernel census of SASS mnemonics in foldseek_b200/libfsb200.so (cuobjdump -sass, sm_100a): which kernels use the DPX min/max
instructions, warp match / redux, 128-bit loads, shared-memory atomics, and that no asynchronous-copy instruction (LDGSTS, UTMALDG,
UBLKCP) or mbarrier (SYNCS) is in the binary.  Run in the build container (no GPU needed):  python tools/sass_census.py > profiles/r02_sass_census.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "foldseek_b200", "libfsb200.so")
KEYS = ["VIADDMNMX", "VIMNMX3", "VIMNMX", "SHFL", "MATCH", "REDUX", "VOTE", "POPC", "LDG.E.128", "LDG.E.64", "LDG.E.U8", "STG.E.128", "LDS.128", "STS.128", "ATOMS", "ATOMG",
        "BAR.SYNC", "LDGSTS", "UTMALDG", "UBLKCP", "SYNCS", "HMMA", "UTCHMMA", "DMUL", "DFMA", "FFMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["c++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out))
    cur, counts, total = None, collections.OrderedDict(), collections.Counter()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        m = re.search(r"/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        counts[cur]["instructions"] += 1
        for k in KEYS:
            if op.startswith(k):
                counts[cur][k] += 1
                break
    print("SASS mnemonic census of foldseek_b200/libfsb200.so (build id %s), sm_100a" % subprocess.run(
        ["python", "-c", "import sys; sys.path.insert(0, %r); import __graft_entry__ as g; print(g.source_id())" % ROOT], capture_output=True, text=True).stdout.strip())
    print("columns: instructions, then the mnemonics that occur; VIADDMNMX / VIMNMX3 = DPX fused add+min/max and three-way min/max\n")
    for fn, c in counts.items():
        name = re.sub(r"\(.*", "", demangle.get(fn, fn))
        name = re.sub(r"^void ", "", name)
        if name.startswith("cub::") or name.startswith("void cub::") or "cub::" in name[:40]:
            name = "cub:: " + re.sub(r"<.*", "", name.split("cub::")[-1])
        rest = "  ".join("%s %d" % (k, c[k]) for k in KEYS if c[k])
        print("%-70s %6d  %s" % (name[:70], c["instructions"], rest))
        total.update(c)
    print("\nwhole library: %d instructions; " % total["instructions"] + ", ".join("%s %d" % (k, total[k]) for k in KEYS))
    print("asynchronous copies (LDGSTS / UTMALDG / UBLKCP) and mbarrier (SYNCS): %d - every global load is a plain LDG into registers"
          % (total["LDGSTS"] + total["UTMALDG"] + total["UBLKCP"] + total["SYNCS"]))


if __name__ == "__main__":
    sys.exit(main())
