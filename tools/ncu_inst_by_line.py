#!/usr/bin/env python
"""Executed warp instructions per source line of one ncu report (needs -lineinfo and
--import-source on): where the instruction budget goes.
usage tools/ncu_inst_by_line.py report.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 50
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source",
                          "cuda,sass"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    cur, hdr, agg = None, None, []
    for r in rows:
        if len(r) >= 2 and r[0] == "File Path":
            cur, hdr = r[1], None
            continue
        if r and r[0] == "Line No":
            hdr = {}
            for i, h in enumerate(r):
                hdr.setdefault(h, i)
            continue
        if not (hdr and cur and len(r) > 8 and r[0].strip().isdigit()):
            continue
        try:
            e = float(r[hdr["Instructions Executed"]] or 0)
        except ValueError:
            continue
        agg.append((cur.split("/")[-1], int(r[0]), r[1].strip()[:90], e))
    tot = sum(a[3] for a in agg) or 1.0
    byfile = collections.Counter()
    for a in agg:
        byfile[a[0]] += a[3]
    print("warp instructions %.4g" % tot)
    for f, s in byfile.most_common(8):
        print("  %-28s %.1f%%" % (f, 100 * s / tot))
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        print("%-24s %4d %5.2f%%  %s" % (a[0], a[1], 100 * a[3] / tot, a[2]))


if __name__ == "__main__":
    main()
