"""SASS opcode histogram of the in-tree library, per kernel family (the evidence behind "tcgen05 / TMEM / TMA" claims:
B200_PROFILING.md names the mnemonics -- UTCHMMA / UTCQMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTCBAR =
tcgen05.commit, UBLKCP = cp.async.bulk (1-D TMA), UTMALDG / UTMASTG = cp.async.bulk.tensor, LDGSTS = cp.async,
REDUX = redux.sync, DADD / DMUL / DSETP = FP64).

    python profiles/sass_histogram.py > profiles/r02_sass_opcodes.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "e2e-mappo-for-mt-fjsp_b200", "libmtfjsp_b200.so")
KEY = ("UTCHMMA", "UTCQMMA", "UTCOMMA", "LDTM", "STTM", "UTCBAR", "UTCATOMSWS", "UTCCP", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS",
       "SYNCS", "REDUX", "DADD", "DMUL", "DFMA", "DSETP", "MUFU", "F2F", "LDS", "STS", "LDG", "STG", "SHFL", "LDL", "STL", "HMMA", "IMMA")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    per = collections.OrderedDict()
    cur = None
    for ln in out.splitlines():
        m = re.search(r"Function : (\S+)", ln)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", name)
            cur = per.setdefault(name, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)", ln)
        if m and cur is not None:
            cur[m.group(1).split(".")[0]] += 1
    tot = collections.Counter()
    print("SASS opcode histogram of %s (cuobjdump -sass), %d kernels\n" % (os.path.basename(LIB), len(per)))
    print("whole library, selected mnemonics:")
    for c in per.values():
        tot.update(c)
    print("  " + "  ".join("%s %d" % (k, tot[k]) for k in KEY if tot[k]))
    print("  (absent: %s)\n" % ", ".join(k for k in KEY if not tot[k]))
    for name, c in per.items():
        n = sum(c.values())
        sel = "  ".join("%s %d" % (k, c[k]) for k in KEY if c[k])
        top = ", ".join("%s %d" % kv for kv in c.most_common(6))
        print("%s\n    %d instructions; %s\n    top: %s" % (name[:200], n, sel, top))


if __name__ == "__main__":
    main()
