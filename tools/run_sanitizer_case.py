#!/usr/bin/env python
"""The case profiles/*_memcheck.txt is taken on: 4096 Mackey-Glass dop853 trajectories (the
headline kernel, dopri_dde1_kernel, plus dopri_init_kernel), then 512 through the general
kernel (dop853 with a critical time), a dopri5 batch, the delayed-SIR difeq batch and the
ring interpolation; results checked against the CPU checker.
usage: compute-sanitizer --tool memcheck python tools/run_sanitizer_case.py"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dde_b200  # noqa: E402
from dde_b200 import workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402

backend = "ref" if pyoracle.available("ref") else "oracle"
model, y0, parms, times, kw = workloads.mackey_glass(4096)
got = dde_b200.dopri_batch(y0, times, model, parms, **kw)
want = pyoracle.dopri_batch(backend, model, y0[:256], times, parms[:256], **kw)
assert np.array_equal(got.statistics[:256], want.statistics)
print("dde1 kernel: 4096 trajectories, %d failed, statistics of the first 256 as the checker's"
      % int((got.code != 1).sum()))
got = dde_b200.dopri_batch(y0[:512], times, model, parms[:512], tcrit=[17.0], **kw)
want = pyoracle.dopri_batch(backend, model, y0[:64], times, parms[:64], tcrit=[17.0], **kw)
assert np.array_equal(got.statistics[:64], want.statistics)
print("general kernel (critical time): 512 trajectories ok")
got = dde_b200.dopri_batch(y0[:512], times, model, parms[:512], **dict(kw, method="dopri5"))
print("dopri5: 512 trajectories, %d failed" % int((got.code != 1).sum()))
b = dde_b200.DopriBatch(model, 1, 512, n_parms=3, **kw)
b.upload(y0[:512], parms[:512])
b.set_times(times)
b.run()
try:
    b.interpolate(np.linspace(460.0, 500.0, 33))
except dde_b200.DdeError as e:  # failed trajectories end early: refused, as R does
    print("interpolate:", str(e)[:60])
b.close()
model, y0, parms, steps, kw = workloads.sir_delay(2048, n_steps=500)
dde_b200.difeq_replicate(2048, y0, steps, model, parms, n_history=16)
print("difeq: 2048 replicates ok")
