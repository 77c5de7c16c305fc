#!/usr/bin/env python
"""Secondary workloads of BASELINE.json (configs[0], [2], [3], [4]) on one GPU.
Not the headline bench (bench.py is configs[1]); prints one JSON line per
workload: device-resident time (the handle's CUDA events), end to end with host
buffers (upload + run + download, wall clock), and the algorithmic rates against
the measured ceilings (FMA-free FP64 issue rate probed live, HBM copy bandwidth
from MEASURED_PEAKS.json) as SURVEY.md section 8d asks for every configuration.

usage: tools/bench_configs.py [--c3-traj N] [--c3-times N] [--c4-rep N] [--c4-steps N]
"""
import argparse
import json
import sys
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import dde_b200  # noqa: E402
from dde_b200 import workloads  # noqa: E402


PEAKS = {}


def peaks():
    """FMA-free FP64 TFLOP/s (probed) and HBM GB/s (MEASURED_PEAKS.json, else the fallback)."""
    if not PEAKS:
        import ctypes as C
        from dde_b200 import _lib
        fn = _lib.load().dde_fp64_peak
        fn.restype = C.c_double
        fn.argtypes = [C.c_int]
        PEAKS["fp64_nofma_tflops"] = fn(0)
        path = ROOT / "MEASURED_PEAKS.json"
        PEAKS["hbm_gbs"] = (float(json.loads(path.read_text())["hbm_gbs"]) if path.exists()
                            else 6650.0)
    return PEAKS


def rates(flops, nbytes, ms):
    pk = peaks()
    tf = flops / (ms * 1e-3) / 1e12
    gbs = nbytes / (ms * 1e-3) / 1e9
    return {"algorithmic_tflops": tf, "fp64_frac_of_nofma": tf / pk["fp64_nofma_tflops"],
            "algorithmic_GBps": gbs, "hbm_frac": gbs / pk["hbm_gbs"]}


def time_dopri(model, y0, parms, times, kw, n, reps=3, e2e=True):
    b = dde_b200.DopriBatch(model, n, y0.shape[0], n_parms=parms.shape[1], **kw)
    b.set_times(times)
    b.upload(y0, parms)
    b.run()
    b.sync()
    ms = []
    for _ in range(reps):
        b.run()
        b.sync()
        ms.append(b.last_ms())
    stats = np.empty((y0.shape[0], 4), dtype=np.int32)
    code = np.empty((y0.shape[0],), dtype=np.int32)
    b.download_into(None, None, stats, code, None, None)
    e2e_ms = None
    if e2e:  # host buffers in, host buffers out (pinned), the streamed download
        import torch
        nt = b.nt
        h_y0 = torch.from_numpy(np.ascontiguousarray(y0)).pin_memory()
        h_p = torch.from_numpy(np.ascontiguousarray(parms)).pin_memory()
        h_y = torch.empty((y0.shape[0], nt, n), dtype=torch.float64).pin_memory()
        h_s = torch.empty((y0.shape[0], 4), dtype=torch.int32).pin_memory()
        h_c = torch.empty((y0.shape[0],), dtype=torch.int32).pin_memory()
        h_t = torch.empty((y0.shape[0],), dtype=torch.float64).pin_memory()
        h_h = torch.empty((y0.shape[0],), dtype=torch.float64).pin_memory()
        t = []
        for _ in range(reps + 1):
            b.sync()
            t0 = time.perf_counter()
            b.upload(h_y0, h_p)
            b.run_download_into(h_y, None, h_s, h_c, h_t, h_h)
            t.append(1e3 * (time.perf_counter() - t0))
        e2e_ms = float(np.median(t[1:]))
    b.close()
    return float(np.median(ms)), stats, code, e2e_ms


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--c3-traj", type=int, default=1 << 22,
                    help="trajectories of EACH of the two C3 sweeps (4 M x 1000 rows = 96 GB of rows)")
    ap.add_argument("--c3-times", type=int, default=1001)
    ap.add_argument("--c4-rep", type=int, default=1 << 20)
    ap.add_argument("--c4-steps", type=int, default=10000)
    ap.add_argument("--c5-traj", type=int, default=4096)
    args = ap.parse_args()

    # C1: one Lorenz trajectory, t = 0..100 (latency, not throughput)
    y0 = np.array([[10.0, 1.0, 1.0]])
    p = np.array([[10.0, 28.0, 8.0 / 3.0]])
    ms, stats, code, e2e_ms = time_dopri("lorenz", y0, p, np.linspace(0, 100, 1001),
                                         dict(method="dopri5"), 3)
    print(json.dumps({"workload": "C1 lorenz dopri5 single trajectory t=0..100", "ms": ms,
                      "e2e_ms": e2e_ms,
                      "statistics": stats[0].tolist(), "code": int(code[0]),
                      "accepted_steps_per_s": stats[0, 2] / (ms * 1e-3)}))

    # C3: Lorenz / Roessler dopri5 sweeps, dense output at c3_times - 1 times
    for gen in (workloads.lorenz_sweep, workloads.rossler_sweep):
        print(json.dumps(c3(gen, args.c3_traj, args.c3_times)))

    print(json.dumps(c4(args.c4_rep, args.c4_steps)))


def c3(gen, n_traj, n_times, e2e=True):
    """C3: one of the two dopri5 parameter sweeps (workloads.lorenz_sweep / rossler_sweep) with
    dense output at n_times - 1 times; returns the result line"""
    model, y0, parms, times, kw = gen(n_traj, n_times=n_times)
    out_bytes = n_traj * (n_times - 1) * 3 * 8
    ms, stats, code, _ = time_dopri(model, y0, parms, times, kw, 3, reps=2, e2e=False)
    S, A, E = (stats[:, k].astype(np.float64).sum() for k in (1, 2, 0))
    frhs = 9 if model == "lorenz" else 7
    flops = 3 * (76 * S + 6 * A + 8 * n_traj * (n_times - 1)) + E * frhs
    line = {"workload": "C3 %s dopri5 sweep" % model, "n_traj": n_traj,
            "n_out_times": n_times - 1, "ms": ms,
            "trajectories_per_s": n_traj / (ms * 1e-3),
            "accepted_steps_per_s": A / (ms * 1e-3),
            "output_GB": out_bytes / 1e9, "failed": int((code != 1).sum()),
            "bound": "FP64 issue (rows leave as whole 32-byte sectors: HBM is not the limit)"}
    line.update(rates(flops, out_bytes + 8.0 * 6 * n_traj, ms))
    if e2e:
        # end to end on a slice whose 6.3 GB of rows fit a pinned host buffer comfortably
        n_e2e = min(n_traj, 1 << 18)
        _, _, _, e2e_ms = time_dopri(model, y0[:n_e2e], parms[:n_e2e], times, kw, 3, reps=2)
        line["e2e"] = {"n_traj": n_e2e, "ms": e2e_ms,
                       "trajectories_per_s": n_e2e / (e2e_ms * 1e-3),
                       "d2h_GBps": n_e2e * (n_times - 1) * 24 / (e2e_ms * 1e-3) / 1e9}
    return line


def c4(n_rep, n_steps):
    """C4: delayed SIR map through difeq_replicate; returns the result line"""
    model, y0, parms, steps, kw = workloads.sir_delay(n_rep, n_steps=n_steps)
    b = dde_b200.DifeqBatch(model, 3, n_rep, n_parms=2, n_history=kw["n_history"])
    b.upload(y0, parms)
    b.set_steps(steps, kw["return_initial"])
    b.run()
    b.sync()
    ms = []
    for _ in range(3):
        b.run()
        b.sync()
        ms.append(b.last_ms())
    ms = float(np.median(ms))
    y, _, code, _ = b.download()
    # end to end: pinned host buffers in, pinned host buffers out (what a caller of the C-ABI
    # hands over, as in the headline's leg)
    import torch
    h_y0 = torch.from_numpy(np.ascontiguousarray(y0)).pin_memory()
    h_p = torch.from_numpy(np.ascontiguousarray(parms)).pin_memory()
    h_y = torch.empty(y.shape, dtype=torch.float64).pin_memory()
    h_c = torch.empty((n_rep,), dtype=torch.int32).pin_memory()
    h_f = torch.empty((n_rep, 3), dtype=torch.float64).pin_memory()
    t = []
    for _ in range(4):
        b.sync()
        t0 = time.perf_counter()
        b.upload(h_y0, h_p)
        b.run()
        b.download_into(h_y, None, h_c, h_f)
        t.append(1e3 * (time.perf_counter() - t0))
    assert np.array_equal(h_y.numpy().view(np.uint64), y.view(np.uint64))
    b.close()
    n_steps_total = float(n_rep) * n_steps
    line = {"workload": "C4 delayed SIR difeq_replicate", "n_rep": n_rep,
            "n_steps": n_steps, "ms": ms, "e2e_ms": float(np.median(t[1:])),
            "steps_per_s": n_steps_total / (ms * 1e-3),
            "replicates_per_s": n_rep / (ms * 1e-3),
            "history": "on chip (shared memory); the ring figure is what an HBM ring would move",
            "bound": "issue latency of the 8-flop recurrence (FP64 fraction of the FMA-free rate)",
            "failed": int((code != 1).sum()), "checksum": float(y.sum())}
    line.update(rates(8.0 * n_steps_total, 40.0 * n_steps_total, ms))
    return line


def c5(n_traj):
    # C5: age-structured SEIR, n = 512, block per trajectory; the ring is
    # n_traj x 128 x 32.8 KB, so one GPU holds ~32k trajectories (65536 shard over 8)
    model, y0, parms, times, kw = workloads.seir_age(n_traj)
    ms, stats, code, e2e_ms = time_dopri(model, y0, parms, times, kw, 512, reps=2)
    S, A, E = (stats[:, k].astype(np.float64).sum() for k in (1, 2, 0))
    n = 512
    frhs = 128 * (2 * 9 * 2 + 22) + 384 * 17  # band sums + rates, lag interpolation
    flops = n * (158 * S + 153 * A + 14 * n_traj * 20) + E * frhs
    ring_bytes = 8.0 * (8 * n + 2) * A + 8.0 * 8 * 384 * E
    line = {"workload": "C5 seir_age dop853 n=512 block-per-trajectory", "n_traj": n_traj,
            "ms": ms, "e2e_ms": e2e_ms, "trajectories_per_s": n_traj / (ms * 1e-3),
            "accepted_steps_per_s": A / (ms * 1e-3), "failed": int((code != 1).sum())}
    line["bound"] = ("latency: the look-ups read their coefficients from L2 at every evaluation, "
                     "block barriers, one thread sums the ordered norms")
    line.update(rates(flops, ring_bytes, ms))
    return line


if __name__ == "__main__":
    t0 = time.time()
    main()
    if "--c5-traj" in " ".join(sys.argv) or True:
        ap = argparse.ArgumentParser()
        ap.add_argument("--c5-traj", type=int, default=4096)
        print(json.dumps(c5(ap.parse_known_args()[0].c5_traj)))
    print("elapsed %.1f s" % (time.time() - t0), file=sys.stderr)
