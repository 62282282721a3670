#!/usr/bin/env python
"""Per-kernel counts of the SASS mnemonics that show which hardware path a kernel uses, from the built objects
(zenslam_b200/_build/*.o, sm_100a).  No GPU needed.

    python tools/sass_summary.py > profiles/sass_summary.txt

UTCIMMA / UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st (TMEM), UTCBAR = tcgen05.commit, UTMALDG = TMA tensor load
(cp.async.bulk.tensor), SYNCS = mbarrier, IDP.2A = dp2a, REDUX = warp reduce, VIMNMX3 = 3-input packed min/max, POPC = popcount,
LDL/STL = local-memory (spill) traffic.
"""
import collections
import glob
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
WATCH = ["UTCIMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTCBAR", "UTMALDG", "UTMASTG", "SYNCS", "IDP.2A", "IDP.4A", "REDUX", "VIMNMX3", "VIMNMX",
         "POPC", "SHFL", "LDS", "STS", "LDG", "STG", "LDL", "STL", "IMAD", "FFMA", "DFMA", "MUFU", "BAR"]


def main():
    objs = sorted(glob.glob(os.path.join(ROOT, "zenslam_b200", "_build", "*.o")))
    if not objs:
        sys.exit("no objects: run python -m zenslam_b200.build first")
    print("# SASS mnemonic counts per kernel (static instruction counts; cuobjdump -sass of zenslam_b200/_build/*.o, sm_100a)")
    print("# columns: total instructions, then every watched mnemonic that occurs")
    for o in objs:
        out = subprocess.run(["cuobjdump", "-sass", o], capture_output=True, text=True).stdout
        kernel, counts, total = None, None, 0
        rows = []
        for line in out.splitlines():
            m = re.match(r"\s*Function : (\S+)", line)
            if m:
                if kernel:
                    rows.append((kernel, total, counts))
                kernel, counts, total = m.group(1), collections.Counter(), 0
                continue
            m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_.]*)", line)
            if m and kernel:
                op = m.group(1)
                total += 1
                for w in WATCH:
                    if op == w or op.startswith(w + "."):
                        counts[w] += 1
                        break
        if kernel:
            rows.append((kernel, total, counts))
        print("\n## %s" % os.path.basename(o))
        for kernel, total, counts in rows:
            name = subprocess.run(["c++filt", kernel], capture_output=True, text=True).stdout.strip() or kernel
            name = re.sub(r"\(.*\)$", "", name)
            body = "  ".join("%s=%d" % (w, counts[w]) for w in WATCH if counts[w])
            print("%-58s %6d  %s" % (name[:58], total, body))


if __name__ == "__main__":
    main()
