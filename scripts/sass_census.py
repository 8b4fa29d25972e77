#!/usr/bin/env python
"""Census of Blackwell-native SASS in mico_b200/lib/libmico_b200.so: per kernel, the number of tcgen05 MMA (UTC*MMA), TMEM
load / store (LDTM / STTM), TMA (UTMALDG / UTMASTG / UBLKCP) and legacy tensor-path (HMMA) instructions.
    python scripts/sass_census.py > profiles/r2_sass_census.txt        (runs without a GPU: cuobjdump -sass)"""
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "mico_b200", "lib", "libmico_b200.so")
PATS = [("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTCHMMA", r"\bUTCHMMA\b(?!\.2CTA)"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"),
        ("UTMALDG", r"\bUTMALDG"), ("UTMASTG", r"\bUTMASTG"), ("UBLKCP", r"\bUBLKCP"), ("UTCBAR", r"\bUTCBAR"),
        ("HMMA(legacy)", r"\bHMMA"), ("MUFU.EX2", r"\bMUFU\.EX2")]


def demangle(name):
    try:
        return subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip() or name
    except OSError:
        return name


def short_name(fn):
    n = demangle(fn).replace("(anonymous namespace)::", "").replace("void ", "").replace("mico::", "")
    return re.sub(r"\(.*", "", n)


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            per[cur] = collections.Counter()
            continue
        if cur is None:
            continue
        per[cur]["instructions"] += 1 if re.search(r"/\*[0-9a-f]{4}\*/", line) else 0
        for k, pat in PATS:
            if re.search(pat, line):
                per[cur][k] += 1
    cols = [k for k, _ in PATS]
    print(f"# SASS census of {os.path.relpath(LIB, REPO)} (cuobjdump -sass, sm_100a); {len(per)} kernels")
    print("# " + " ".join(f"{c:>12}" for c in ["instr"] + cols) + "  kernel")
    tot = collections.Counter()
    for fn, c in per.items():
        tot.update(c)
        if not any(c[k] for k in cols[:-1]):
            continue
        short = short_name(fn)
        print("  " + " ".join(f"{c[k]:>12}" for k in ["instructions"] + cols) + "  " + short[:110])
    print("# total " + " ".join(f"{k}={tot[k]}" for k in cols) + f" instructions={tot['instructions']}")
    simt = [short_name(fn) for fn, c in per.items() if not any(c[k] for k in cols[:-1])]
    print("# SIMT-only kernels (HBM-bound helpers): " + ", ".join(sorted(set(simt))))


if __name__ == "__main__":
    sys.exit(main())
