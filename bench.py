#!/usr/bin/env python
"""bench.py -- headline benchmark: Mackey-Glass DDE trajectories/sec.

Workload (BASELINE.json configs[1]): 1,048,576
Mackey-Glass trajectories (tau = 17) per GPU, random (beta, gamma, n, y0) from
the fixed splitmix64 stream, dop853, rtol = atol = 1e-6, n_history = 64, 11
output times on [0, 500].  One "step" = one pass of the hot path over that
batch (every trajectory integrated from t = 0 to t = 500).

  value  trajectories/s, whole job, inputs already resident in HBM, timed with
         CUDA events on the launching stream, max over ranks.
  e2e    the same metric through the reference-facing C-ABI with HOST
         (pinned) buffers: per step H2D of y0/parms, then
         dde_dopri_batch_run_download (the solve with the D2H of y_out,
         statistics, codes, t_last, step_size streamed block by block behind
         it), wall clock between device synchronisations.
  --impl reference   the reference's own CPU code (oracle/_ref: unmodified
         src/dopri*.c, one forked process per host core) on a bounded sample
         of the same workload.

Multi-GPU: trajectories are independent, so ranks take disjoint slices of the
parameter stream (weak scaling, no collective on the data path); torch
distributed is used only for the barrier and the max-over-ranks of the time.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

N_TRAJ_PER_GPU = 1 << 20
F_POW = 43          # FP64 instructions of dde_pow's main path
F_RHS_MG = 6 + 17 + F_POW  # model 6, lag theta + 853 Horner 3 + 14, pow


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--n-traj", type=int, default=N_TRAJ_PER_GPU, help="trajectories per GPU")
    ap.add_argument("--cpu-sample", type=int, default=0,
                    help="trajectories in the CPU baseline sample (0 = auto)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-secondary", action="store_true",
                    help="skip the C4 / C5 extras and the strong-scaling leg")
    return ap.parse_args()


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML from a thread
    (every 10 ms; an nvidia-smi process per rank is too slow to answer inside a
    sub-second region on an 8-GPU box), nvidia-smi as the fallback."""

    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []
        self.sm, self.smax, self.reasons = [], [], set()
        self.nvml = None
        self.stop_flag = threading.Event()

    def start(self):
        try:
            import pynvml
            pynvml.nvmlInit()
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            self.nvml = pynvml
            self.thread = threading.Thread(target=self._poll, daemon=True)
            self.thread.start()
            return
        except Exception:
            self.nvml = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.FIELDS,
                 "--format=csv,noheader,nounits", "-lms", "100"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _poll(self):
        nv = self.nvml
        masks = []
        for name, attr in (("hw_slowdown", "nvmlClocksEventReasonHwSlowdown"),
                           ("hw_thermal_slowdown", "nvmlClocksEventReasonHwThermalSlowdown"),
                           ("sw_thermal_slowdown", "nvmlClocksEventReasonSwThermalSlowdown"),
                           ("sw_power_cap", "nvmlClocksEventReasonSwPowerCap")):
            bit = getattr(nv, attr, None)
            if bit is None:
                bit = getattr(nv, attr.replace("ClocksEventReason", "ClocksThrottleReason"), None)
            if bit is not None:
                masks.append((name, bit))
        get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or \
            getattr(nv, "nvmlDeviceGetCurrentClocksThrottleReasons", None)
        try:
            smax = float(nv.nvmlDeviceGetMaxClockInfo(self.handle, nv.NVML_CLOCK_SM))
        except Exception:
            smax = None
        while not self.stop_flag.is_set():
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.handle, nv.NVML_CLOCK_SM)))
                if smax is not None:
                    self.smax.append(smax)
                if get_reasons is not None:
                    r = get_reasons(self.handle)
                    for name, bit in masks:
                        if r & bit:
                            self.reasons.add(name)
            except Exception:
                pass
            self.stop_flag.wait(0.01)

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.nvml is not None:
            self.stop_flag.set()
            self.thread.join(timeout=2)
            return self._summary(self.sm, self.smax, self.reasons, "nvml")
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                smax.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        return self._summary(sm, smax, reasons, "nvidia-smi")

    @staticmethod
    def _summary(sm, smax, reasons, source):
        # "under load": the upper half of the samples (idle gaps between steps drop the clock)
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": float(np.median(load)) if load else None,
                "sm_max_mhz": max(smax) if smax else None, "reasons": sorted(reasons),
                "samples": len(sm), "source": source}


def algorithmic_work(stats, n=1, n_parms=3, n_lags=1, n_lag_idx=1, nt=11, order=8):
    """Per-batch algorithmic FLOPs and bytes from the per-trajectory statistics
    (SURVEY.md section 8d, dop853 row)."""
    E = stats[:, 0].astype(np.float64)
    S = stats[:, 1].astype(np.float64)
    A = stats[:, 2].astype(np.float64)
    T = float(nt)
    flops = n * (158.0 * S + 153.0 * A + 14.0 * T) + 3.0 * T + E * F_RHS_MG + S * (F_POW + 12.0)
    hist = 8.0 * (order * n + 2) * A + 8.0 * (order * n_lag_idx + 2) * S * n_lags
    bytes_ = 8.0 * (n + n_parms) + 8.0 * n * T + hist + 4.0 * 5 + 8.0 * 2
    return float(flops.sum()), float(bytes_.sum())


def source_hash():
    """sha256 over the device sources (include/dde_b200, dde_b200/csrc, dde_b200/models): what a committed
    ncu capture is valid for"""
    import hashlib
    h = hashlib.sha256()
    for d in ("include/dde_b200", "dde_b200/csrc", "dde_b200/models"):
        for f in sorted((ROOT / d).iterdir()):
            if f.is_file():
                h.update(f.name.encode())
                h.update(f.read_bytes())
    return h.hexdigest()


def strong_scaling_leg(torch, dist, args, world, rank, park, weak):
    """The 1 M-trajectory batch through the in-process sharded handle (dde_dopri_shards_*:
    what the R glue would call in place of the serial loop of src/r_difeq.c:94-103), host
    buffers to host buffers, over 1 device and over all `world` devices of the box.  Rank 0
    alone runs it, the other ranks wait on a CPU (gloo) barrier, their GPUs idle."""
    import dde_b200
    from dde_b200 import workloads
    if rank != 0:
        dist.barrier(group=park)
        return None
    out = None
    try:
        n_total = N_TRAJ_PER_GPU
        n_dev = torch.cuda.device_count()
        if n_dev < world:
            raise RuntimeError("rank 0 sees %d devices, %d needed" % (n_dev, world))
        model, y0, parms, times, kw = workloads.mackey_glass(n_total, first=0)
        nt = times.shape[0]
        h_y0 = torch.from_numpy(y0).pin_memory()
        h_parms = torch.from_numpy(parms).pin_memory()
        h_y = torch.empty((n_total, nt, 1), dtype=torch.float64).pin_memory()
        h_stats = torch.empty((n_total, 4), dtype=torch.int32).pin_memory()
        h_code = torch.empty((n_total,), dtype=torch.int32).pin_memory()
        h_t = torch.empty((n_total,), dtype=torch.float64).pin_memory()
        h_h = torch.empty((n_total,), dtype=torch.float64).pin_memory()
        res = {}
        for n_shards in sorted({1, world}):
            sh = dde_b200.DopriShards(model, 1, n_total, devices=list(range(n_shards)), n_parms=3,
                                      **kw)
            sh.set_times(times)
            dev_ms, wall = [], []
            for i in range(args.warmup + args.steps):
                t0 = time.perf_counter()
                sh.upload(h_y0, h_parms)
                sh.run_download_into(h_y, None, h_stats, h_code, h_t, h_h)  # blocking
                if i >= args.warmup:
                    wall.append(time.perf_counter() - t0)
                    dev_ms.append(sh.last_ms())
            res[n_shards] = {"value": n_total * len(dev_ms) / (float(np.sum(dev_ms)) * 1e-3),
                             "e2e": n_total * len(wall) / float(np.sum(wall)),
                             "ms_per_step": float(np.mean(dev_ms)),
                             "e2e_ms_per_step": 1e3 * float(np.mean(wall))}
            # the same batch run AGAIN, trajectories handed out longest first by the step
            # counts of the run before (dde_dopri_shards_order_by_last_run): what a handle
            # that repeats a batch can do about the tail; reported beside, never as, the value
            sh.order_by_last_run(True)
            dev_ms, wall = [], []
            for i in range(1 + args.steps):
                t0 = time.perf_counter()
                sh.upload(h_y0, h_parms)
                sh.run_download_into(h_y, None, h_stats, h_code, h_t, h_h)
                if i >= 1:
                    wall.append(time.perf_counter() - t0)
                    dev_ms.append(sh.last_ms())
            res[n_shards]["rerun_longest_first"] = {
                "value": n_total * len(dev_ms) / (float(np.sum(dev_ms)) * 1e-3),
                "e2e": n_total * len(wall) / float(np.sum(wall)),
                "ms_per_step": float(np.mean(dev_ms))}
            sh.close()
        same = None
        if weak is not None and weak[0].shape[0] == n_total:
            # rank 0's own weak-scaling batch is this batch: every bit must agree
            same = bool(np.array_equal(h_stats.numpy(), weak[0]) and
                        np.array_equal(h_code.numpy(), weak[1]) and
                        np.array_equal(h_y.numpy().view(np.uint64), weak[2].view(np.uint64)))
        top = res[world]
        out = {"n_traj_total": n_total, "n_gpus": world, "value": top["value"], "e2e": top["e2e"],
               "unit": "trajectories/s", "ms_per_step": top["ms_per_step"],
               "e2e_ms_per_step": top["e2e_ms_per_step"],
               "n1": res[1],
               "efficiency_vs_n1": top["value"] / (world * res[1]["value"]),
               "e2e_efficiency_vs_n1": top["e2e"] / (world * res[1]["e2e"]),
               "bits_equal_rank_local_result": same,
               "rerun_longest_first": {
                   "value": top["rerun_longest_first"]["value"],
                   "e2e": top["rerun_longest_first"]["e2e"],
                   "ms_per_step": top["rerun_longest_first"]["ms_per_step"],
                   "efficiency_vs_n1": top["rerun_longest_first"]["value"] /
                   (world * res[1]["rerun_longest_first"]["value"]),
                   "efficiency_vs_n1_index_order": top["rerun_longest_first"]["value"] /
                   (world * res[1]["value"]),
                   "what": "the batch run again with trajectories handed out longest first by "
                           "the step counts of the run before (exact costs: an upper bound on "
                           "what any hand-out order can do; a 1 % perturbation of the "
                           "parameters between the runs already loses most of it)"},
               "how": "dde_dopri_shards_* in rank 0's process, one host thread per device, pinned "
                      "host buffers in and out; value = trajectories / slowest shard's device "
                      "time, e2e = wall clock around upload + run_download",
               "note": "granularity-bound: one device runs 148 SMs x 512 = 75776 trajectories at "
                       "a time, so the batch is 13.8 trajectories per resident thread on 1 GPU "
                       "and 1.7 on 8; it ends when the slowest thread's last trajectory does "
                       "(about one trajectory's duration, 2.8 ms, after the device stops being "
                       "full), which is 7 % of the run on 1 GPU and most of it on 8.  Measured "
                       "(profiles/r2_bench_*gpu.json): 0.93 / 0.81 / 0.64 at 2 / 4 / 8 GPUs; the "
                       "weak-scaling value above is linear."}
    except Exception as exc:  # the leg must never take the headline line with it
        out = {"unavailable": "%s: %s" % (type(exc).__name__, exc)}
    if park is not None:
        dist.barrier(group=park)
    return out


def cpu_reference_run(n_sample, first, n_proc):
    """Time the unmodified reference on `n_sample` trajectories of the workload."""
    from dde_b200 import workloads
    from oracle import pyoracle
    backend = "ref" if pyoracle.available("ref") else "oracle"
    model, y0, parms, times, kw = workloads.mackey_glass(n_sample, first=first)
    r = pyoracle.dopri_batch(backend, model, y0, times, parms, n_proc=n_proc if backend == "ref"
                             else 0, want_y=True, **kw)
    return backend, r


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = len(os.sched_getaffinity(0))
    # ~1.5-3 k trajectories/s/core: keep each step to a few seconds
    n_sample = args.cpu_sample or max(2048, 1024 * cores)
    times_s = []
    backend = "ref"
    for i in range(args.warmup + args.steps):
        backend, r = cpu_reference_run(n_sample, first=0, n_proc=cores)
        if i >= args.warmup:
            times_s.append(r.seconds)
    total = float(np.sum(times_s))
    value = n_sample * len(times_s) / total
    kind = "reference" if backend == "ref" else "port"
    sample = ("first %d trajectories of the 1,048,576-trajectory Mackey-Glass batch per step, "
              "%s, %d forked processes" % (n_sample, "unmodified reference C (oracle/_ref)"
                                           if backend == "ref" else "oracle port", cores))
    line = {
        "impl": "reference", "metric": "Mackey-Glass DDE trajectories/sec", "value": value,
        "unit": "trajectories/s", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * total / len(times_s), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "mackey_glass_dop853_tau17 (BASELINE.json configs[1])",
                   "n_traj_per_step": n_sample, "rtol": 1e-6, "atol": 1e-6, "n_history": 64,
                   "times": "0,50,...,500"},
        "cpu_baseline": {"value": value, "unit": "trajectories/s", "cores": cores, "kind": kind,
                         "sample": sample},
        "e2e": {"value": value, "unit": "trajectories/s", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    args = parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
        return

    import torch
    import torch.distributed as dist

    import dde_b200
    from dde_b200 import workloads

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: dde_b200 has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local_rank)
    park = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
        park = dist.new_group(backend="gloo")  # a barrier that keeps the GPUs idle

    n_traj = args.n_traj
    model, y0, parms, times, kw = workloads.mackey_glass(n_traj, first=rank * n_traj)
    nt = times.shape[0]

    # pinned host buffers: what a caller of the C-ABI hands over / gets back
    h_y0 = torch.from_numpy(y0).pin_memory()
    h_parms = torch.from_numpy(parms).pin_memory()
    h_y = torch.empty((n_traj, nt, 1), dtype=torch.float64).pin_memory()
    h_stats = torch.empty((n_traj, 4), dtype=torch.int32).pin_memory()
    h_code = torch.empty((n_traj,), dtype=torch.int32).pin_memory()
    h_t = torch.empty((n_traj,), dtype=torch.float64).pin_memory()
    h_h = torch.empty((n_traj,), dtype=torch.float64).pin_memory()
    h2d = h_y0.numel() * 8 + h_parms.numel() * 8
    d2h = h_y.numel() * 8 + h_stats.numel() * 4 + h_code.numel() * 4 + h_t.numel() * 8 + h_h.numel() * 8

    b = dde_b200.DopriBatch(model, 1, n_traj, n_parms=3, **kw)
    b.set_times(times)
    b.upload(h_y0, h_parms)
    b.sync()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # ---- device-resident: warm-up, then K timed steps -------------------------
    for _ in range(args.warmup):
        b.run()
    b.sync()
    sampler = ClockSampler(local_rank)
    barrier()
    sampler.start()
    kernel_ms = []
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.run()
        b.sync()
        kernel_ms.append(b.last_ms())  # CUDA events on the handle's stream
    t_wall = time.perf_counter() - t0
    launches = args.steps * b.last_launches()
    barrier()
    dev_ms = float(np.sum(kernel_ms))
    t_dev = torch.tensor([dev_ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_dev, op=dist.ReduceOp.MAX)
    dev_ms_max = float(t_dev.item())

    # ---- end to end through the C-ABI with host buffers ------------------------
    for _ in range(1):
        b.upload(h_y0, h_parms)
        b.run_download_into(h_y, None, h_stats, h_code, h_t, h_h)
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        b.upload(h_y0, h_parms)
        b.run_download_into(h_y, None, h_stats, h_code, h_t, h_h)  # blocking
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    clocks = sampler.stop()
    t_e2e = torch.tensor([e2e_s], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_s_max = float(t_e2e.item())

    # ---- the same batch run again, handed out longest first by the last run's step counts
    #      (an extra: the size of the tail in which lanes have run out of work) ----
    rerun = None
    if not args.no_secondary:
        b.order_by_last_run()
        b.run()
        b.sync()
        barrier()
        rr_ms = []
        for _ in range(args.steps):
            b.run()
            b.sync()
            rr_ms.append(b.last_ms())
        b.set_order(None)
        t_rr = torch.tensor([float(np.sum(rr_ms))], dtype=torch.float64, device="cuda")
        if world > 1:
            dist.all_reduce(t_rr, op=dist.ReduceOp.MAX)
        rerun = float(t_rr.item())

    stats = h_stats.numpy()
    code = h_code.numpy()
    acc = torch.tensor([float(stats[:, 2].sum())], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(acc, op=dist.ReduceOp.SUM)

    # ---- strong scaling of the 1 M batch through the in-process sharded handle ----
    strong = None
    if not args.no_secondary:
        barrier()
        weak = (stats.copy(), code.copy(), h_y.numpy().copy()) if rank == 0 else None
        strong = strong_scaling_leg(torch, dist, args, world, rank, park, weak)
        barrier()

    if rank == 0:
        total_traj = n_traj * world
        value = total_traj * args.steps / (dev_ms_max * 1e-3)
        e2e_value = total_traj * args.steps / e2e_s_max
        flops, bytes_ = algorithmic_work(stats, nt=nt)
        ms_kernel = dev_ms / args.steps
        peaks_path = ROOT / "MEASURED_PEAKS.json"
        if peaks_path.exists():
            hbm_peak = float(json.loads(peaks_path.read_text())["hbm_gbs"])
            peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)"
        else:
            hbm_peak = 6650.0
            peak_src = "fallback (B200_PROFILING.md)"
        achieved_gbs = bytes_ / (ms_kernel * 1e-3) / 1e9
        # dram__bytes_read.sum + dram__bytes_write.sum of the stepping kernel from the
        # committed ncu --set full capture of this command (profiles/traffic.json): used only
        # if it was taken at this batch size from exactly these device sources
        traffic = None
        ncu_pipe = None
        traffic_note = "no capture committed for these device sources"
        tpath = ROOT / "profiles" / "traffic.json"
        if tpath.exists():
            tj = json.loads(tpath.read_text())
            if tj.get("n_traj_per_gpu") == n_traj and tj.get("source_sha256") == source_hash():
                traffic = tj.get("dram_bytes_per_launch")
                traffic_note = tj.get("source")
                # the same capture's pipe counters: the flop convention counts a division
                # or a pow as written, the pipe executes their expansions
                ncu_pipe = {k: tj[k] for k in ("fp64_pipe_pct_of_peak", "issue_active_pct",
                                               "threads_per_warp_instruction") if k in tj}
        fp64 = measure_fp64_peaks(torch)
        achieved_tflops = flops / (ms_kernel * 1e-3) / 1e12
        # the evaluation the reference makes at the old time of every accepted dop853 step
        # (src/dopri.c:400-403) is counted in n_eval but not computed by this kernel
        skipped = float(stats[:, 2].astype(np.float64).sum()) * F_RHS_MG
        line = {
            "metric": "Mackey-Glass DDE trajectories/sec", "value": value, "unit": "trajectories/s",
            "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {
                "workload": "mackey_glass_dop853_tau17 (BASELINE.json configs[1])",
                "n_traj_per_gpu": n_traj, "rtol": 1e-6, "atol": 1e-6, "n_history": 64,
                "times": "0,50,...,500", "parallelism": "replicate-sharded x%d" % world,
                "l2": "working set per step (%.1f GB of history rings) exceeds the 126 MB L2"
                      % (n_traj * 64 * 10 * 8 / 1e9)},
            "accepted_steps_per_s_per_gpu": float(acc.item()) / world * args.steps /
                                            (dev_ms_max * 1e-3),
            "failed_trajectories": int((code != 1).sum()),
            "roofline": {
                "bound": "fp64_nofma", "achieved": achieved_tflops, "peak": fp64["nofma"],
                "unit": "TFLOP/s",
                "frac": achieved_tflops / fp64["nofma"] if fp64["nofma"] else None,
                "traffic": traffic, "traffic_source": traffic_note,
                "peak_source": "measured live (dde_fp64_peak: alternating DMUL / DADD chains, the "
                               "issue rate FMA-free code can reach; DFMA: %s TFLOP/s)"
                               % ("%.1f" % fp64["fma"] if fp64["fma"] else "n/a"),
                "flops_convention": "SURVEY.md 8d per trajectory: 158 S + 153 A + 14 T + 3 T + "
                                    "E (6 + 17 + F_pow) + S (F_pow + 12), one + - * / fmax as "
                                    "written = 1 flop, F_pow = %d, from the run's own statistics "
                                    "(S attempts, A accepts, E evaluations, T rows); ONE pow per "
                                    "attempt (8d lists two: dop853 has step_beta = 0 and its "
                                    "pow(fac_old, 0) is not evaluated); E counts the reference's "
                                    "redundant evaluation per accepted step, which this kernel "
                                    "counts and does not compute" % F_POW,
                "achieved_without_uncomputed_evaluations":
                    (flops - skipped) / (ms_kernel * 1e-3) / 1e12,
                "ncu": ncu_pipe,
                "hbm": {"achieved": achieved_gbs, "peak": hbm_peak, "unit": "GB/s",
                        "frac": achieved_gbs / hbm_peak, "peak_source": peak_src,
                        "note": "algorithmic bytes per trajectory: inputs 32, rows 88, statistics "
                                "36, ring commits 80 A, lag reads 80 S"}},
            "e2e": {"value": e2e_value, "unit": "trajectories/s", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h},
            "gpu_launches": launches,
            "clocks": clocks,
            "host_wall_ms_per_step": 1e3 * t_wall / args.steps,
        }
        if rerun is not None:
            line["rerun_longest_first"] = {
                "value": world * n_traj * args.steps / (rerun * 1e-3), "unit": "trajectories/s",
                "ms_per_step": rerun / args.steps,
                "roofline_frac": (flops / (rerun / args.steps * 1e-3) / 1e12 / fp64["nofma"]
                                  if fp64["nofma"] else None),
                "what": "NOT the headline: the same batch run again with trajectories handed "
                        "out longest first by the step counts of the run before "
                        "(dde_dopri_batch_order_by_last_run) -- exact costs, so this is what "
                        "the tail of the batch, in which lanes have run out of work, costs the "
                        "value above; results are bit-identical"}
        if strong is not None:
            line["strong_1M"] = strong
        if world == 1 and not args.no_secondary:
            # the two secondary configurations cheap enough for every run (tools/bench_configs.py
            # has all of them): C4 at its full size, C5 at 4096 trajectories
            try:
                sys.path.insert(0, str(ROOT / "tools"))
                import bench_configs
                from dde_b200 import workloads as _w
                # C3 at a quarter of its size per sweep (1 M x 1000 rows = 24 GB of rows on the
                # device; the full 4 M per sweep: tools/bench_configs.py, profiles/), kernel only
                # -- end to end these are bound by moving the rows over PCIe
                line["c3_lorenz_1M"] = bench_configs.c3(_w.lorenz_sweep, 1 << 20, 1001, e2e=False)
                line["c3_rossler_1M"] = bench_configs.c3(_w.rossler_sweep, 1 << 20, 1001, e2e=False)
                line["c4_difeq_sir_delay"] = bench_configs.c4(1 << 20, 10000)
                line["c5_seir_age_4096"] = bench_configs.c5(4096)
            except Exception as exc:
                line["secondary_unavailable"] = "%s: %s" % (type(exc).__name__, exc)
        if world == 1 and not args.no_cpu_baseline:
            cores = len(os.sched_getaffinity(0))
            n1 = args.cpu_sample or 8192
            backend, r1 = cpu_reference_run(n1, 0, 0)
            nN = max(n1, 4096 * cores)
            _, rN = cpu_reference_run(nN, 0, cores)
            # parity spot check on the sample the baseline just integrated
            same = bool(np.array_equal(r1.statistics, stats[:n1]) and
                        np.array_equal(r1.code, code[:n1]))
            line["cpu_baseline"] = {
                "value": nN / rN.seconds, "unit": "trajectories/s", "cores": cores,
                "kind": "reference" if backend == "ref" else "port",
                "sample": "first %d trajectories of the batch over %d forked processes "
                          "(unmodified reference C, gcc -O2); single core: first %d trajectories"
                          % (nN, cores, n1),
                "single_core_value": n1 / r1.seconds,
                "gpu_matches_cpu_statistics_on_sample": same}
        print(json.dumps(line))
    b.close()
    if world > 1:
        dist.destroy_process_group()


def measure_fp64_peaks(torch):
    """FP64 issue ceilings measured live: register-resident dependent-free
    chains of DFMA (2 flop/instr) and of alternating DMUL/DADD (1 flop/instr,
    the ceiling that applies to FMA-free code)."""
    try:
        from dde_b200 import _lib
        lib = _lib.load()
        import ctypes as C
        fn = lib.dde_fp64_peak
        fn.restype = C.c_double
        fn.argtypes = [C.c_int]
        return {"fma": fn(1), "nofma": fn(0)}
    except Exception:
        return {"fma": None, "nofma": None}


if __name__ == "__main__":
    main()
