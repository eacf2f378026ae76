#!/usr/bin/env python
"""Static instruction counts per kernel of libfdms.so (cuobjdump -sass): the mnemonics that show what a kernel is built
from -- UBLKCP (cp.async.bulk, the TMA bulk copy), SYNCS (mbarrier), IDP (8-bit dot products), vector loads / stores,
shared memory, atomics, MUFU, fp64, warp shuffles / votes, barriers.
Usage: python tools/sass_summary.py > profiles/<name>.txt"""
import collections
import os
import re
import subprocess
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from kernel_resources import kernel_name  # noqa: E402

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "face_detection_multi_scale_b200", "libfdms.so")


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur:
            cnt[cur][m.group(1)] += 1
    rows = []
    for f, c in cnt.items():
        dem = subprocess.run(["cu++filt", f], capture_output=True, text=True).stdout.strip() or f
        dem = kernel_name(re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", dem))

        def pre(*ps):
            return sum(v for k, v in c.items() if k.startswith(ps))
        rows.append((dem, sum(c.values()), pre("UBLKCP"), pre("SYNCS"), pre("IDP"), pre("LDG"), pre("STG"),
                     pre("LDS", "STS"), pre("ATOM", "RED"), pre("MUFU"), pre("DFMA", "DMUL", "DADD", "DSETP"),
                     pre("SHFL", "VOTE"), pre("BAR")))
    print("# cuobjdump -sass face_detection_multi_scale_b200/libfdms.so: static instruction counts per kernel (sm_100a)")
    print("# UBLKCP = cp.async.bulk (TMA bulk copy), SYNCS = mbarrier operations, IDP = 8-bit dot products (IDP.2A / IDP.4A)")
    print("%-40s %6s %6s %5s %4s %4s %4s %7s %8s %4s %4s %9s %3s" % (
        "kernel", "instrs", "UBLKCP", "SYNCS", "IDP", "LDG", "STG", "LDS+STS", "ATOM+RED", "MUFU", "FP64", "SHFL+VOTE", "BAR"))
    for r in sorted(rows):
        print("%-40s %6d %6d %5d %4d %4d %4d %7d %8d %4d %4d %9d %3d" % r)


if __name__ == "__main__":
    main()
