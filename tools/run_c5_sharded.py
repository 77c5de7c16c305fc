#!/usr/bin/env python
"""C5 at BASELINE.json's full size: 65 536 age-structured SEIR trajectories (n = 512,
dop853, lags 14 and 21, 128-entry rings of 32.8 KB entries = 275 GB of history) sharded
over all GPUs of the box in ONE process (dde_dopri_batch_sharded: one host thread and
stream per device, no collective).  Host buffers in, host buffers out; wall clock.
Prints one JSON line.  usage: tools/run_c5_sharded.py [n_traj]"""
import json
import sys
import time
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np  # noqa: E402
import torch  # noqa: E402

import dde_b200  # noqa: E402
from dde_b200 import workloads  # noqa: E402

n_traj = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
n_dev = torch.cuda.device_count()
model, y0, parms, times, kw = workloads.seir_age(n_traj)
wall = []
for _ in range(2):  # the first call also creates the contexts
    t0 = time.perf_counter()
    r = dde_b200.dopri_batch(y0, times, model, parms, devices="all", **kw)
    wall.append(time.perf_counter() - t0)
# the same trajectories alone on one device: a shard's bits do not depend on the sharding
k = 64
one = dde_b200.dopri_batch(y0[-k:], times, model, parms[-k:], **kw)
same = bool(np.array_equal(one.y.view(np.uint64), r.y[-k:].view(np.uint64)) and
            np.array_equal(one.statistics, r.statistics[-k:]))
print(json.dumps({
    "workload": "C5 seir_age dop853 n=512, %d trajectories sharded over %d GPUs in one process"
                % (n_traj, n_dev),
    "n_gpus": n_dev, "n_traj": n_traj, "wall_s_first_call": wall[0], "wall_s": wall[1],
    "trajectories_per_s_e2e": n_traj / wall[1],
    "accepted_steps_per_s_e2e": float(r.statistics[:, 2].astype(np.float64).sum()) / wall[1],
    "failed": int((r.code != 1).sum()),
    "last_%d_bit_identical_to_single_device_run" % k: same,
}))
