#!/usr/bin/env python
"""profiles/r02_sass_summary.txt: counts of the Blackwell-specific SASS instructions per kernel of libmdt_b200.so (cuobjdump -sass; runs without a GPU).
usage: python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "medicaldetectiontoolkit_b200", "libmdt_b200.so")
PATS = [("UTCHMMA", r"\bUTCHMMA"), ("UTCBAR", r"\bUTCBAR"), ("LDTM", r"\bLDTM"), ("UTMALDG", r"\bUTMALDG"), ("SYNCS", r"\bSYNCS"), ("REDG", r"\bREDG"),
        ("ELECT", r"\bELECT"), ("LDGSTS", r"\bLDGSTS"), ("R2UR.BC", r"R2UR\.BROADCAST")]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
    counts = collections.OrderedDict()
    cur = None
    it = iter(names)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = re.sub(r"\(.*", "", next(it))
            counts[cur] = collections.Counter()
            continue
        if cur is None or "/*" not in line:
            continue
        if re.search(r"/\*[0-9a-f]{4,}\*/", line):
            counts[cur]["instrs"] += 1
            for key, pat in PATS:
                if re.search(pat, line):
                    counts[cur][key] += 1
    print("SASS summary of medicaldetectiontoolkit_b200/libmdt_b200.so (cuobjdump -sass, sm_100a), end of round 2.  Blackwell-specific instructions per kernel:")
    print("UTCHMMA = tcgen05.mma, UTCBAR = tcgen05.commit, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (TMA), SYNCS = mbarrier, REDG = red.global,")
    print("ELECT = elect.sync, LDGSTS = cp.async, R2UR.BC = per-instruction uniform-register waterfalls (0 = issue loops run from one elected thread).\n")
    print("%-64s %7s" % ("kernel", "instrs") + "".join(" %8s" % k for k, _ in PATS))
    for name, c in counts.items():
        print("%-64s %7d" % (name[:64], c["instrs"]) + "".join(" %8d" % c[k] for k, _ in PATS))
    tot = collections.Counter()
    for c in counts.values():
        tot.update(c)
    print("%-64s %7d" % ("total (%d kernels)" % len(counts), tot["instrs"]) + "".join(" %8d" % tot[k] for k, _ in PATS))


if __name__ == "__main__":
    main()
