#!/usr/bin/env python
"""Per-source-line stall samples of one ncu report (needs -lineinfo and --import-source on):
usage tools/ncu_lines.py report.ncu-rep [top]"""
import collections
import csv
import io
import subprocess
import sys


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
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
            s = float(r[hdr["# Samples"]] or 0)
            e = float(r[hdr["Instructions Executed"]] or 0)
        except ValueError:
            continue
        stalls = {k[6:]: float(r[i] or 0) for k, i in hdr.items()
                  if k.startswith("stall_") and "Not Issued" not in k}
        agg.append((cur.split("/")[-1], int(r[0]), r[1].strip()[:84], s, e, stalls))
    tot = sum(a[3] for a in agg) or 1.0
    byfile = collections.Counter()
    for a in agg:
        byfile[a[0]] += a[3]
    print("samples %d" % tot)
    for f, s in byfile.most_common(8):
        print("  %-28s %.1f%%" % (f, 100 * s / tot))
    for a in sorted(agg, key=lambda a: -a[3])[:top]:
        st = sorted(a[5].items(), key=lambda kv: -kv[1])[:2]
        print("%-22s %4d %5.1f%%  %-22s %s" % (a[0], a[1], 100 * a[3] / tot,
              " ".join("%s %.0f%%" % (k, 100 * v / max(a[3], 1)) for k, v in st), a[2]))


if __name__ == "__main__":
    main()
