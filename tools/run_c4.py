#!/usr/bin/env python
"""C4 only (delayed SIR difeq_replicate), for profiling: usage tools/run_c4.py [n_rep] [n_steps]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import numpy as np
import dde_b200
from dde_b200 import workloads
n_rep = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
n_steps = int(sys.argv[2]) if len(sys.argv) > 2 else 2000
model, y0, parms, steps, kw = workloads.sir_delay(n_rep, n_steps=n_steps)
b = dde_b200.DifeqBatch(model, 3, n_rep, n_parms=2, n_history=kw["n_history"])
b.upload(y0, parms)
b.set_steps(steps, kw["return_initial"])
for _ in range(2):
    b.run()
    b.sync()
print("ms", b.last_ms(), "steps/s %.3e" % (n_rep * n_steps / (b.last_ms() * 1e-3)))
b.close()
