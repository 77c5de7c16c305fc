#!/usr/bin/env python
"""Debug: lag-window hit/miss counters of the stats build (variants/libdde_stats.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DDE_B200_LIB"] = os.path.abspath("variants/libdde_stats.so")
import numpy as np
import dde_b200
from dde_b200 import workloads, _lib
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model, y0, parms, times, kw = workloads.mackey_glass(n)
lib = _lib.load()
out = (C.c_ulonglong * 8)()
lib.dde_debug_lag_stats(out)
r = dde_b200.dopri_batch(y0, times, model, parms, **kw)
lib.dde_debug_lag_stats(out)
names = ["miss_above", "miss_below", "miss_empty", "ensure_appends", "failures", "ring_probes", "answered_from_ring", "above_but_predicted"]
st = r.statistics.sum(axis=0)
print("per trajectory: evals %.1f steps %.1f accepts %.1f" % tuple(st[:3] / n))
for k, v in zip(names, out):
    print("%-16s %10d  %.2f per trajectory" % (k, v, v / n))
