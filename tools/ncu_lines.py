#!/usr/bin/env python
"""Per-source-line totals from an .ncu-rep captured with --import-source on:
instructions executed, share, average active threads, stall samples.
Usage: python tools/ncu_lines.py x.ncu-rep [min_share_percent]"""
import csv
import subprocess
import sys


def main():
    rep = sys.argv[1]
    min_share = float(sys.argv[2]) if len(sys.argv) > 2 else 0.8
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    fname = "?"
    lines = []
    hdr = None
    for r in rows:
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Line No":
            hdr = r
            ie = hdr.index("Instructions Executed")
            ti = hdr.index("Thread Instructions Executed")
            ns = hdr.index("# Samples")
            continue
        if hdr is None or r[0] in ("Function Name",) or r[0] == "":
            continue
        try:
            lines.append((int(r[ie]), int(r[ti]), int(r[ns]), fname, r[0], r[1].strip()))
        except (ValueError, IndexError):
            pass
    tot = sum(x[0] for x in lines)
    samples = sum(x[2] for x in lines)
    print("total warp instructions %d, samples %d" % (tot, samples))
    for v, t, s, f, ln, src in sorted(lines, key=lambda x: (x[3], int(x[4]))):
        if v >= tot * min_share / 100 or s >= samples * min_share / 100:
            print("%5.2f%% inst %5.2f%% samp  thr %4.1f  %s:%s  %s" % (100.0 * v / tot, 100.0 * s / max(1, samples),
                                                                     t / max(1, v), f, ln, src[:90]))


if __name__ == "__main__":
    main()
