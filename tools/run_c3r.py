#!/usr/bin/env python
"""C3 Roessler half only: usage tools/run_c3r.py [n_traj] [n_times]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import dde_b200
from dde_b200 import workloads
n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
n_times = int(sys.argv[2]) if len(sys.argv) > 2 else 1001
model, y0, parms, times, kw = workloads.rossler_sweep(n_traj, n_times=n_times)
b = dde_b200.DopriBatch(model, 3, n_traj, n_parms=3, **kw)
b.set_times(times)
b.upload(y0, parms)
for _ in range(2):
    b.run()
    b.sync()
print("rossler ms", b.last_ms(), "traj/s %.4g" % (n_traj / (b.last_ms() * 1e-3)))
b.close()
