#!/usr/bin/env python
"""Developer aid: per-source-line executed instructions and stall samples from an .ncu-rep (needs -lineinfo)."""
import csv
import subprocess
import sys


def main(rep, top=40, by="inst"):
    out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"],
                         capture_output=True, text=True).stdout
    cur, agg = None, {}
    for r in csv.reader(out.splitlines()):
        if not r:
            continue
        if r[0] == 'File Path':
            cur = r[1].split('/')[-1]
            continue
        if r[0] in ('Function Name', 'Line No'):
            continue
        if r[0] != '' and len(r) > 7:
            try:
                n, s = int(r[7]), int(r[4])
            except ValueError:
                continue
            a = agg.setdefault((cur, r[0]), [0, 0, r[1][:105]])
            a[0] += n
            a[1] += s
    ti = sum(a[0] for a in agg.values())
    ts = sum(a[1] for a in agg.values())
    print("total inst %d  samples %d" % (ti, ts))
    key = (lambda kv: -kv[1][0]) if by == "inst" else (lambda kv: -kv[1][1])
    for k, a in sorted(agg.items(), key=key)[:top]:
        print("inst %10d %5.1f%%  samp %6d %5.1f%%  %s:%s %s" % (a[0], 100 * a[0] / ti, a[1], 100 * a[1] / max(ts, 1), k[0], k[1], a[2]))


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40, sys.argv[3] if len(sys.argv) > 3 else "inst")
