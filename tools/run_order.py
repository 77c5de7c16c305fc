#!/usr/bin/env python
"""The headline batch with and without the hand-out order of its last run: usage
tools/run_order.py [n_traj] [perturb] -- perturb: relative size of a perturbation of every
parameter between the run the order is taken from and the run that is timed"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import dde_b200
from dde_b200 import workloads
n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
eps = float(sys.argv[2]) if len(sys.argv) > 2 else 0.0
model, y0, parms, times, kw = workloads.mackey_glass(n_traj)
b = dde_b200.DopriBatch(model, 1, n_traj, n_parms=3, **kw)
b.set_times(times)
b.upload(y0, parms)


def timed(tag, reps=3):
    ms = []
    for _ in range(reps):
        b.run()
        b.sync()
        ms.append(b.last_ms())
    print("%-34s %.3f ms  %.4g traj/s" % (tag, np.median(ms), n_traj / (np.median(ms) * 1e-3)))


timed("index order")
if eps > 0:
    rng = np.random.default_rng(3)
    b.upload(y0, parms * (1 + eps * rng.uniform(-1, 1, parms.shape)))
    b.run()
    b.sync()
b.order_by_last_run()
b.upload(y0, parms)
timed("longest first (last run, eps %g)" % eps)
b.set_order(None)
timed("index order again")
b.close()
