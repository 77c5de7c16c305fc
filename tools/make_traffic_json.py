#!/usr/bin/env python
"""profiles/traffic.json from one `ncu --set full` capture of the stepping kernel:
dram bytes per launch and the pipe / issue / lane counters bench.py quotes, keyed on the
batch size AND on a hash of the device sources the capture was taken from (bench.py emits
`traffic: null` when either differs, so a kernel change cannot leave a stale figure behind).
usage: tools/make_traffic_json.py report.ncu-rep n_traj_per_gpu"""
import csv
import json
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
import bench  # noqa: E402


def main():
    rep, n_traj = sys.argv[1], int(sys.argv[2])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True,
                         text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d, u = dict(zip(hdr, vals)), dict(zip(hdr, units))

    def nbytes(key):
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}[u[key]]
        return float(d[key].replace(",", "")) * scale

    rd, wr = nbytes("dram__bytes_read.sum"), nbytes("dram__bytes_write.sum")
    tj = {
        "n_traj_per_gpu": n_traj,
        "source_sha256": bench.source_hash(),
        "dram_bytes_per_launch": rd + wr, "dram_bytes_read": rd, "dram_bytes_write": wr,
        "kernel": d.get("Kernel Name", ""),
        "kernel_ms": float(d["gpu__time_duration.sum"].replace(",", "")),
        "source": "%s (ncu --set full --clock-control none, second launch of `bench.py --steps 1 "
                  "--warmup 1 --n-traj %d`, dram__bytes_read.sum + dram__bytes_write.sum)"
                  % (Path(rep).name, n_traj),
        "fp64_pipe_pct_of_peak":
            float(d["sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active"]),
        "issue_active_pct": float(d["smsp__issue_active.avg.pct_of_peak_sustained_active"]),
        "threads_per_warp_instruction":
            float(d["smsp__thread_inst_executed_per_inst_executed.ratio"]),
    }
    (ROOT / "profiles" / "traffic.json").write_text(json.dumps(tj, indent=1) + "\n")
    print(json.dumps(tj, indent=1))


if __name__ == "__main__":
    main()
