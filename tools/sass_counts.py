#!/usr/bin/env python
"""Per-kernel SASS instruction counts of opentn_b200/libotn.so (cuobjdump -sass): which kernels carry DMMA.8x8x4, cp.async (LDGSTS),
bulk copies (UBLKCP), mbarrier operations (SYNCS) ...   usage: tools/sass_counts.py > profiles/sass_counts_xxx.csv"""
import collections, os, re, subprocess, sys

LIB = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "opentn_b200", "libotn.so")
COLS = ["DMMA", "LDGSTS", "UBLKCP", "SYNCS", "DFMA", "DMUL", "DADD", "LDG", "STG", "LDS", "STS", "BAR", "SHFL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    counts, order, cur = {}, [], None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
            counts[cur] = collections.Counter()
            order.append(cur)
            continue
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            counts[cur][m.group(1)] += 1
            counts[cur]["total"] += 1
    dm = demangle(order)
    print("# per-kernel SASS instruction counts of libotn.so (cuobjdump -sass, sm_100a); tools/sass_counts.py")
    print("kernel,total," + ",".join(COLS))
    for k in sorted(order, key=lambda k: dm[k]):
        name = re.sub(r"\(.*", "", dm[k]).replace("void ", "")
        print('"%s",%d,%s' % (name, counts[k]["total"], ",".join(str(counts[k][c]) for c in COLS)))


if __name__ == "__main__":
    main()
