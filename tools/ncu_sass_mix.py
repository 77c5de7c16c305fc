#!/usr/bin/env python
"""Instruction mix and stall samples per SASS opcode from an ncu report's source page.
usage: tools/ncu_sass_mix.py report.ncu-rep [top_n]"""
import csv
import subprocess
import sys
from collections import defaultdict


def main():
    path = sys.argv[1]
    top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
    out = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[1]
    ix = {k: i for i, k in enumerate(hdr)}
    inst = defaultdict(float)
    thr = defaultdict(float)
    samp = defaultdict(float)
    stall_cols = [k for k in hdr if k.startswith("stall_") and "Not Issued" not in k]
    stall = defaultdict(lambda: defaultdict(float))
    tot_i = tot_s = tot_t = 0.0
    for r in rows[2:]:
        if len(r) < len(hdr):
            continue
        sass = r[ix["Source"]].split()
        if not sass:
            continue
        op = sass[0]
        if op.startswith("@"):
            op = sass[1]
        op = op.split(".")[0]
        n = float(r[ix["Instructions Executed"]] or 0)
        t = float(r[ix["Thread Instructions Executed"]] or 0)
        s = float(r[ix["# Samples"]] or 0)
        inst[op] += n
        thr[op] += t
        samp[op] += s
        tot_i += n
        tot_s += s
        tot_t += t
        for k in stall_cols:
            stall[op][k] += float(r[ix[k]] or 0)
    print("total warp inst %.3e  thread inst %.3e (avg %.1f/32)  samples %d" %
          (tot_i, tot_t, tot_t / tot_i, tot_s))
    print("%-10s %8s %8s %6s   top stalls" % ("op", "inst%", "samp%", "thr"))
    for op, n in sorted(inst.items(), key=lambda kv: -samp[kv[0]])[:top]:
        st = sorted(stall[op].items(), key=lambda kv: -kv[1])[:3]
        print("%-10s %7.2f%% %7.2f%% %6.1f   %s" %
              (op, 100 * n / tot_i, 100 * samp[op] / tot_s, thr[op] / n if n else 0,
               ", ".join("%s %.0f" % (k[6:], v) for k, v in st if v > 0)))


if __name__ == "__main__":
    main()
