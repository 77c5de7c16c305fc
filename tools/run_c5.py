#!/usr/bin/env python
"""C5 only (age-structured SEIR, n = 512, block per trajectory), for profiling:
usage tools/run_c5.py [n_traj]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import dde_b200
from dde_b200 import workloads
n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
model, y0, parms, times, kw = workloads.seir_age(n_traj)
b = dde_b200.DopriBatch(model, 512, n_traj, n_parms=3, **kw)
b.set_times(times)
b.upload(y0, parms)
for _ in range(2):
    b.run()
    b.sync()
r = b.download(want_y=False)
print("ms", b.last_ms(), "traj/s %.4g" % (n_traj / (b.last_ms() * 1e-3)), "accepted steps/s %.4g" %
      (r.statistics[:, 2].sum() / (b.last_ms() * 1e-3)), "failed", int((r.code != 1).sum()))
b.close()
