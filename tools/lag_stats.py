#!/usr/bin/env python
"""Debug: lag-window hit/miss counters of the stats build (variants/libdde_stats.so)."""
import ctypes as C, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["DDE_B200_LIB"] = os.path.abspath("variants/libdde_stats.so")
import numpy as np
import dde_b200
from dde_b200 import workloads, _lib
dde_b200.load_model_library("tests/models/libdde_test_models.so")  # seir, delayed_growth
n = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
model, y0, parms, times, kw = workloads.mackey_glass(n)
lib = _lib.load()
out = (C.c_ulonglong * 16)()
lib.dde_debug_lag_stats(out)
r = dde_b200.dopri_batch(y0, times, model, parms, **kw)
lib.dde_debug_lag_stats(out)
names = ["miss_above", "miss_below", "miss_empty", "ensure_appends", "failures", "ring_probes", "answered_from_ring"]
st = r.statistics.sum(axis=0)
print("per trajectory: evals %.1f steps %.1f accepts %.1f" % tuple(st[:3] / n))
for k, v in zip(names, out):
    print("%-16s %10d  %.2f per trajectory" % (k, v, v / n))
print("index check violations: %d" % out[15])
# the stress cases of tests/test_parity_gpu.py under the checking build
viol = 0
for n_history in (3, 5, 8, 16, 33, 64):
    model, y0, parms, times, kw = workloads.mackey_glass(2048, first=2048)
    kw["n_history"] = n_history
    dde_b200.dopri_batch(y0, times, model, parms, **kw)
    lib.dde_debug_lag_stats(out)
    viol += out[15]
for method, rtol in (("dopri5", 1e-6), ("dopri853", 1e-9), ("dopri853", 1e-3)):
    model, y0, parms, times, kw = workloads.mackey_glass(1024, first=9000)
    kw.update(method=method, rtol=rtol, atol=rtol, n_history=200)
    dde_b200.dopri_batch(y0, np.linspace(0, 120, 481), model, parms, tcrit=[17.0, 34.0], **kw)
    lib.dde_debug_lag_stats(out)
    viol += out[15]
y0 = np.tile([1e7 - 1, 0.0, 1.0, 0.0], (6, 1))
for method in ("dopri5", "dopri853"):
    dde_b200.dopri_batch(y0, np.linspace(0, 400, 41), "seir", None, n_out=1, method=method,
                         n_history=40)
    lib.dde_debug_lag_stats(out)
    viol += out[15]
rng = np.random.default_rng(11)
dde_b200.dopri_batch(rng.uniform(0.5, 2, (5, 2)), -np.linspace(0, 3, 13), "delayed_growth",
                     rng.uniform(0.1, 1, (5, 2)), n_history=50)
lib.dde_debug_lag_stats(out)
viol += out[15]
print("index check violations over the stress cases: %d" % viol)
