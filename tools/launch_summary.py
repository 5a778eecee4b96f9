#!/usr/bin/env python
"""Compact view of an `ncu --metrics gpu__time_duration.sum --csv` launch list:
kernel (short name), launches, total / mean duration.  Usage: launch_summary.py x.csv"""
import csv
import re
import sys
from collections import OrderedDict


def main():
    rows = [r for r in csv.reader(open(sys.argv[1], errors="replace")) if len(r) > 5]
    hdr = next(r for r in rows if "Kernel Name" in r)
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = OrderedDict()
    for r in rows[rows.index(hdr) + 1:]:
        if len(r) <= vi:
            continue
        name = re.sub(r"\(.*", "", r[ki])
        name = re.sub(r"<.*", "<>", name)[:70]
        val = float(r[vi].replace(",", ""))
        scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}.get(r[ui], 1.0)
        a = agg.setdefault(name, [0, 0.0])
        a[0] += 1
        a[1] += val * scale
    total = sum(a[1] for a in agg.values())
    print("%-70s %8s %12s %12s %7s" % ("kernel", "launches", "total_us", "mean_us", "share"))
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%-70s %8d %12.1f %12.1f %6.1f%%" % (k, n, t, t / n, 100 * t / total))


if __name__ == "__main__":
    main()
