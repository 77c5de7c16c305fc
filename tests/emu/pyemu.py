"""TEST INFRASTRUCTURE ONLY: ctypes access to tests/emu/libdde_emu.so, the device
code of include/dde_b200 compiled by g++ and run one lane at a time (cuda_emu.h)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

from oracle import pyoracle

HERE = Path(__file__).resolve().parent
_libs = {}


def build(tag="", extra=""):
    import os
    env = dict(os.environ)
    out = "libdde_emu%s.so" % tag
    env["DDE_EMU_OUT"] = out
    env["DDE_EMU_EXTRA"] = extra
    subprocess.run(["bash", str(HERE / "build.sh")], check=True, env=env, capture_output=True)
    return HERE / out


def load(tag="", extra=""):
    key = (tag, extra)
    if key not in _libs:
        path = HERE / ("libdde_emu%s.so" % tag)
        src_time = max(p.stat().st_mtime for p in list(HERE.glob("*.cpp")) + list(HERE.glob("*.h")) +
                       list((HERE.parent.parent / "include" / "dde_b200").glob("*")) +
                       list((HERE.parent.parent / "dde_b200" / "models").glob("*")))
        if not path.exists() or path.stat().st_mtime < src_time:
            build(tag, extra)
        lib = C.CDLL(str(path))
        lib.emu_dopri_batch.restype = C.c_int
        lib.emu_dopri_batch.argtypes = pyoracle._DOPRI_ARGS
        lib.emu_last_message.restype = C.c_char_p
        lib.emu_set_mode.argtypes = [C.c_int]
        _libs[key] = lib
    return _libs[key]


def dopri_batch(model, y, times, parms=None, *, n_out=0, tcrit=None, mode=-1, tag="", extra="",
                **ctl_kwargs):
    lib = load(tag, extra)
    ctl = pyoracle.control(**ctl_kwargs)
    y, p, n_traj, n, n_parms, stride = pyoracle._prep(y, parms)
    times = np.ascontiguousarray(times, dtype=np.float64)
    tc = None if tcrit is None else np.ascontiguousarray(tcrit, dtype=np.float64)
    nt = times.shape[0] if ctl.return_initial else times.shape[0] - 1
    r = pyoracle.Result()
    r.y = np.zeros((n_traj, nt, n))
    r.output = np.zeros((n_traj, nt, max(n_out, 1)))
    r.statistics = np.zeros((n_traj, 4), dtype=np.int32)
    r.code = np.zeros((n_traj,), dtype=np.int32)
    r.t_last = np.zeros((n_traj,))
    r.step_size = np.zeros((n_traj,))
    dp, ip = pyoracle._dp, pyoracle._ip
    lib.emu_set_mode(mode)
    rc = lib.emu_dopri_batch(model.encode(), C.byref(ctl), n, n_out, n_traj, dp(y), dp(p), n_parms,
                             stride, dp(times), times.shape[0], dp(tc),
                             0 if tc is None else tc.shape[0], dp(r.y), dp(r.output),
                             ip(r.statistics), ip(r.code), dp(r.t_last), dp(r.step_size))
    if rc != 0:
        raise RuntimeError(lib.emu_last_message().decode())
    if n_out == 0:
        r.output = None
    return r
