"""Per-kernel SASS mnemonic counts of libub200.so (cuobjdump -sass): which kernels are tcgen05 / TMA / TMEM,
and — for the peer exchange — which carry system-scope release / acquire accesses.

    python tools/sass_summary.py > profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "uniter_b200", "lib", "libub200.so")
COLS = [("UTCHMMA", r"\bUTCHMMA(?!\.2CTA)"), ("UTCHMMA.2CTA", r"\bUTCHMMA\.2CTA"), ("UTMALDG", r"\bUTMALDG"),
        ("UTMASTG", r"\bUTMASTG"), ("LDTM", r"\bLDTM"), ("STTM", r"\bSTTM"), ("UTCBAR", r"\bUTCBAR"),
        ("MUFU.EX2", r"MUFU\.EX2"), ("MUFU.RCP", r"MUFU\.RCP"), ("MUFU.TANH", r"MUFU\.TANH"),
        ("STG.SYS", r"\bSTG\.E(\.\w+)*\.STRONG\.SYS"), ("LDG.SYS", r"\bLDG\.E(\.\w+)*\.STRONG\.SYS"),
        ("MEMBAR.SYS", r"MEMBAR\.\w+\.SYS")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    demangle = {}
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts.setdefault(cur, collections.Counter())["__n"] += 1
            continue
        if cur is None:
            continue
        for name, pat in COLS:
            if re.search(pat, line):
                counts[cur][name] += 1
    names = list(counts)
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    agg = collections.OrderedDict()
    for mangled, dem in zip(names, out):
        base = re.sub(r"<.*", "", re.sub(r"^void ", "", dem))
        base = re.sub(r"\(.*", "", base)
        a = agg.setdefault(base, collections.Counter())
        a.update(counts[mangled])
    print("# SASS summary of uniter_b200/lib/libub200.so (cuobjdump -sass, sm_100a), per kernel template (all instantiations summed)")
    print("# columns: instantiations | " + " | ".join(n for n, _ in COLS))
    for base, c in sorted(agg.items(), key=lambda kv: -(kv[1]["UTCHMMA"] + kv[1]["UTCHMMA.2CTA"] + kv[1]["UTMALDG"])):
        print("%-40s %4d | " % (base[:40], c["__n"]) + " | ".join("%5d" % c[n] for n, _ in COLS))


if __name__ == "__main__":
    sys.exit(main())
