#!/usr/bin/env python
"""Generate tests/golden/*.npz from the UNMODIFIED reference solver.

Run in the build container, where /root/reference exists and oracle/Makefile
has compiled it into oracle/_ref/libdde_ref.so (`make -C oracle ref`).  Each
case stores the inputs and the reference's outputs (results, statistics,
return codes, final time / step size) bit for bit; tests/test_golden_*.py
replay the inputs through the oracle restatement (CPU) and the CUDA path (GPU)
and compare.  The .npz files are committed: the GPU box has no reference tree.
"""
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from dde_b200 import workloads  # noqa: E402
from oracle import pyoracle  # noqa: E402

OUT = ROOT / "tests" / "golden"


def dopri_case(name, model, y0, times, parms, n_out=0, tcrit=None, **kw):
    r = pyoracle.dopri_batch("ref", model, y0, times, parms, n_out=n_out, tcrit=tcrit, **kw)
    np.savez_compressed(
        OUT / ("dopri_%s.npz" % name), kind="dopri", model=model,
        y0=np.atleast_2d(np.asarray(y0, dtype=np.float64)),
        parms=np.zeros((0,)) if parms is None else np.asarray(parms, dtype=np.float64),
        has_parms=parms is not None, times=np.asarray(times, dtype=np.float64),
        tcrit=np.zeros((0,)) if tcrit is None else np.asarray(tcrit, dtype=np.float64),
        n_out=n_out, kw_keys=np.array(sorted(kw)), kw_vals=np.array([str(kw[k]) for k in sorted(kw)]),
        y=r.y, output=np.zeros((0,)) if r.output is None else r.output, statistics=r.statistics,
        code=r.code, t_last=r.t_last, step_size=r.step_size)
    print("dopri_%s: %d trajectories, codes %s" % (name, r.code.shape[0],
                                                   dict(zip(*np.unique(r.code, return_counts=True)))))


def difeq_case(name, model, y0, steps, parms, n_out=0, n_history=0, return_initial=True):
    r = pyoracle.difeq_batch("ref", model, y0, steps, parms, n_out=n_out, n_history=n_history,
                             return_initial=return_initial)
    np.savez_compressed(
        OUT / ("difeq_%s.npz" % name), kind="difeq", model=model,
        y0=np.atleast_2d(np.asarray(y0, dtype=np.float64)),
        parms=np.zeros((0,)) if parms is None else np.asarray(parms, dtype=np.float64),
        has_parms=parms is not None, steps=np.asarray(steps, dtype=np.uint64), n_out=n_out,
        n_history=n_history, return_initial=return_initial, y=r.y,
        output=np.zeros((0,)) if r.output is None else r.output, code=r.code, y_final=r.y_final)
    print("difeq_%s: %d replicates" % (name, r.code.shape[0]))


def main():
    assert pyoracle.available("ref"), "build oracle/_ref first: make -C oracle ref"
    OUT.mkdir(parents=True, exist_ok=True)
    lor_p = [10.0, 28.0, 8.0 / 3.0]
    # BASELINE.json configs[0]: 4565 steps / 4365 acc / 200 rej / 27392 evals
    dopri_case("c1_lorenz", "lorenz", [[10.0, 1.0, 1.0]], np.linspace(0, 100, 1001), lor_p)
    dopri_case("lorenz_output_853", "lorenz", [[10.0, 1.0, 1.0], [1.0, 1.0, 1.0]],
               np.linspace(0, 5, 51), lor_p, n_out=2, method="dopri853")
    # BASELINE.json configs[1]: the first 256 trajectories of the Mackey-Glass batch
    model, y0, parms, times, kw = workloads.mackey_glass(256)
    dopri_case("c2_mackey_glass", model, y0, times, parms, **kw)
    dopri_case("mackey_glass_kat", "mackey_glass", [[1.2]], np.linspace(0, 500, 11),
               [0.2, 0.1, 10.0], method="dopri853", n_history=1000)
    # BASELINE.json configs[2], reduced
    for gen, nm in ((workloads.lorenz_sweep, "c3_lorenz_sweep"),
                    (workloads.rossler_sweep, "c3_rossler_sweep")):
        model, y0, parms, times, kw = gen(64, n_times=101)
        dopri_case(nm, model, y0, times, parms, **kw)
    seir_y0 = [[1e7 - 1, 0.0, 1.0, 0.0]]
    dopri_case("seir_dopri5", "seir", seir_y0, np.linspace(0, 30, 301), None, n_out=1,
               method="dopri5", rtol=1e-7, atol=1e-7, n_history=1000)
    dopri_case("seir_int_853", "seir_int", seir_y0, np.linspace(0, 200, 301), None, n_out=1,
               method="dopri853", rtol=1e-7, atol=1e-7, n_history=1000)
    dopri_case("seir_underflow", "seir", seir_y0, np.linspace(0, 30, 301), None, rtol=1e-7,
               atol=1e-7, n_history=2)
    dopri_case("tcrit", "lorenz", [[10.0, 1.0, 1.0]], np.linspace(0, 2, 21), lor_p,
               tcrit=[0.55, 1.0, 1.25, 7.0])
    dopri_case("backward_dopri5", "exponential", [[1.0, 2.0]], -np.linspace(0, 2, 9),
               [[0.5, 0.25]], method="dopri5")
    dopri_case("backward_853", "exponential", [[1.0, 2.0]], -np.linspace(0, 2, 9), [[0.5, 0.25]],
               method="dopri853")
    dopri_case("stiff", "exponential", [[1.0]], [0.0, 10.0], [[1e5]], stiff_check=1)
    dopri_case("delayed_growth", "delayed_growth", [[1.0, 2.0, 0.5]], np.linspace(0, 3, 13),
               [[0.5, 0.25, 1.0]], method="dopri853", n_history=200)
    # BASELINE.json configs[4], reduced: two trajectories of the n = 512 model
    model, y0, parms, times, kw = workloads.seir_age(2)
    dopri_case("c5_seir_age", model, y0, times, parms, **kw)
    rng = np.random.default_rng(70)
    dopri_case("delayed_growth_wide", "delayed_growth_wide", rng.uniform(0.5, 2.0, (3, 70)),
               np.linspace(0, 3, 13), rng.uniform(0.1, 1.0, (3, 70)), method="dopri5",
               n_history=200)
    # difeq
    difeq_case("logistic", "logistic", [[0.1]], np.arange(21), [[1.5]])
    difeq_case("fibonacci", "fibonacci", [[1.0] * 5], np.arange(11), None, n_history=2)
    difeq_case("fib_out", "fib_out", [[1.0]], np.arange(11), None, n_out=1, n_history=2)
    difeq_case("additive_out", "additive", [[1.0, 2.0, 3.0]], [2, 5, 6, 11], [[1.0, 2.0, 3.0]],
               n_out=2, n_history=8, return_initial=False)
    # BASELINE.json configs[3], reduced
    model, y0, parms, steps, kw = workloads.sir_delay(64, n_steps=2000)
    difeq_case("c4_sir_delay", model, y0, steps, parms, **kw)
    model, y0, parms, steps, kw = workloads.sir_delay(4, n_steps=100)
    difeq_case("history_miss", model, y0, steps, parms, n_history=8, return_initial=False)


if __name__ == "__main__":
    main()
