"""Debug: run 128 Mackey-Glass trajectories under the tracing build (variants/libdde_trace.so,
DDE_NVCC_EXTRA="-DDDE_LAG_STATS -DDDE_LAG_TRACE_PRINT"), which prints one line per lag-window miss."""
import os, sys
sys.path.insert(0, os.getcwd())
os.environ["DDE_B200_LIB"] = os.path.abspath("variants/libdde_trace.so")
import dde_b200
from dde_b200 import workloads
model, y0, parms, times, kw = workloads.mackey_glass(128)
r = dde_b200.dopri_batch(y0, times, model, parms, **kw)
print("done", r.statistics.sum(axis=0))
