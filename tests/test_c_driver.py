"""The C-ABI from plain C (examples/c_driver.c), the way the reference's own glue
would call it: compiles against include/dde_b200.h with gcc, links the library,
and gives the numbers the Python binding gives."""
import re
import subprocess
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent


def build_driver(tmp_path):
    exe = tmp_path / "c_driver"
    lib_dir = ROOT / "dde_b200"
    subprocess.run(["gcc", "-std=c99", "-O2", "-Wall", "-Werror", "-I", str(ROOT / "include"),
                    str(ROOT / "examples" / "c_driver.c"), "-L", str(lib_dir), "-ldde_b200",
                    "-Wl,-rpath," + str(lib_dir), "-lm", "-o", str(exe)], check=True)
    return exe


def test_c_driver_builds_and_fails_loudly_without_a_device(tmp_path):
    import torch
    exe = build_driver(tmp_path)
    if torch.cuda.is_available():
        pytest.skip("covered by the GPU test")
    r = subprocess.run([str(exe), "8"], capture_output=True, text=True)
    assert r.returncode == 2
    assert "no CPU fallback" in r.stderr


@pytest.mark.gpu
def test_c_driver_matches_the_python_binding(tmp_path):
    import dde_b200
    from dde_b200 import workloads
    n_traj = 4096
    exe = build_driver(tmp_path)
    r = subprocess.run([str(exe), str(n_traj)], capture_output=True, text=True, check=True)
    m = re.search(r"batch n_traj (\d+) evals (\d+) accepts (\d+) failed (\d+) sum_y_end (\S+)",
                  r.stdout)
    assert m, r.stdout
    model, y0, parms, times, kw = workloads.mackey_glass(n_traj)
    got = dde_b200.dopri_batch(y0, times, model, parms, **kw)
    ok = got.code == 1
    assert int(m.group(1)) == n_traj
    assert int(m.group(2)) == int(got.statistics[:, 0].astype(np.int64).sum())
    assert int(m.group(3)) == int(got.statistics[:, 2].astype(np.int64).sum())
    assert int(m.group(4)) == int((~ok).sum())
    total = 0.0
    for v in got.y[ok, -1, 0]:  # the driver's summation order
        total += float(v)
    assert float(m.group(5)) == total
    if (~ok).any():
        assert "first failure: trajectory %d:" % int(np.flatnonzero(~ok)[0]) in r.stdout
    # the continued handle: run to 250, continue to 500
    m2 = re.search(r"continue completed (\d+) sum_y_end (\S+)", r.stdout)
    assert m2, r.stdout
    b = dde_b200.DopriBatch(model, 1, n_traj, n_parms=3, **kw)
    b.upload(y0, parms)
    b.set_times([0.0, 250.0])
    b.run()
    b.continue_([250.0, 500.0])
    res = b.download()
    b.close()
    ok2 = res.code == 1
    assert int(m2.group(1)) == int(ok2.sum())
    total = 0.0
    for v in res.y[ok2, -1, 0]:
        total += float(v)
    assert float(m2.group(2)) == total
