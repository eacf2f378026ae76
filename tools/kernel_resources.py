#!/usr/bin/env python
"""Registers / stack / shared memory of every kernel in libfdms.so (cuobjdump -res-usage), as a table.
Usage: python tools/kernel_resources.py > profiles/<name>.txt"""
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "face_detection_multi_scale_b200", "libfdms.so")


def kernel_name(dem):
    """`void ns::k<T, (bool)0>(Params)` -> `k<T, (bool)0>`: cut the return type and the parameter list."""
    dem = re.sub(r"^void ", "", dem)
    i = dem.find("<")
    j = dem.find("(")
    if i < 0 or (0 <= j < i):
        return dem[:j] if j >= 0 else dem
    depth = 0
    for k in range(i, len(dem)):
        depth += dem[k] == "<"
        depth -= dem[k] == ">"
        if depth == 0:
            return dem[:k + 1]
    return dem


def main():
    out = subprocess.run(["cuobjdump", "-res-usage", LIB], capture_output=True, text=True).stdout
    rows, name = [], None
    for line in out.splitlines():
        m = re.match(r"\s*Function (\S+):", line)
        if m:
            name = m.group(1)
            continue
        m = re.search(r"REG:(\d+) STACK:(\d+) SHARED:(\d+) LOCAL:(\d+)", line)
        if m and name:
            dem = subprocess.run(["cu++filt", name], capture_output=True, text=True).stdout.strip() or name
            dem = re.sub(r"\(anonymous namespace\)::|<unnamed>::", "", dem)
            dem = kernel_name(dem)
            rows.append((dem, int(m.group(1)), int(m.group(2)), int(m.group(3))))
            name = None
    print("# cuobjdump -res-usage face_detection_multi_scale_b200/libfdms.so (sm_100a)")
    print("%-58s %5s %6s %8s" % ("kernel", "regs", "stack", "smem(B)"))
    for r in sorted(rows):
        print("%-58s %5d %6d %8d" % r)


if __name__ == "__main__":
    main()
