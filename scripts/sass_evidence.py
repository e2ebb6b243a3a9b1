"""Static evidence for the built library (no GPU needed): per kernel, the register / stack / shared
usage `cuobjdump -res-usage` reports and how many Blackwell-native SASS instructions it holds.

    python scripts/sass_evidence.py > profiles/r02_sass_evidence.txt

Mnemonics (B200_PROFILING.md, "What proves a Blackwell-native kernel"): UTC*MMA = tcgen05.mma,
LDTM/STTM = tcgen05.ld/st, UBLKCP = cp.async.bulk (1-D bulk TMA; the operand images here are linear,
so no tensor map is needed), UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = the legacy path
(must be 0).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "human-3d-reconstruction_b200", "libsmpl_b200.so")
COUNTED = ["UTC.MMA", "UTCBAR", "LDTM", "STTM", "UBLKCP", "UTMALDG", "SYNCS", "HMMA", "FFMA", "LDG", "STG"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))


def short(sig):
    sig = re.sub(r"^void ", "", sig)
    sig = re.sub(r"^smplb200::", "", sig)
    return sig.split("(")[0]


def main():
    res = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True, check=True).stdout
    usage = {}
    fn = None
    for line in res.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            fn = m.group(1)
            continue
        if fn and "REG:" in line:
            usage[fn] = dict(kv.split(":") for kv in line.split() if ":" in kv)
            fn = None
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    counts = collections.defaultdict(collections.Counter)
    fn = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        if fn is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if not m:
            continue
        op = m.group(1)
        for pat in COUNTED:
            if re.match(pat, op.split(".")[0]) or (pat == "UTC.MMA" and re.match(r"UTC[A-Z]+MMA", op)):
                counts[fn][pat] += 1
                break
    names = demangle(sorted(usage))
    print(f"# {os.path.relpath(LIB, ROOT)} -- cuobjdump -res-usage + -sass, sm_100a; counts are static SASS instructions")
    head = f"{'kernel':58s} {'REG':>4s} {'STACK':>5s} {'SHARED':>6s} " + " ".join(f"{c:>7s}" for c in COUNTED)
    print(head)
    for fn in sorted(usage, key=lambda f: short(names[f])):
        u = usage[fn]
        row = f"{short(names[fn])[:58]:58s} {u['REG']:>4s} {u['STACK']:>5s} {u['SHARED']:>6s} "
        row += " ".join(f"{counts[fn][c]:7d}" for c in COUNTED)
        print(row)
    legacy = sum(c["HMMA"] for c in counts.values())
    print(f"\n# legacy HMMA instructions in the library: {legacy}")
    return 0


if __name__ == "__main__":
    sys.exit(main())
