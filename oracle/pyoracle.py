"""TEST INFRASTRUCTURE ONLY: ctypes access to the checkers.

  backend "ref"    -> oracle/_ref/libdde_ref.so, the UNMODIFIED reference solver
                      compiled from /root/reference (oracle/Makefile `ref`)
  backend "oracle" -> oracle/liboracle.so, the plain-C restatement
                      (oracle/dde_oracle.c)

Both export the same batch entry points with the argument list of the
product's dde_dopri_batch / dde_difeq_batch (include/dde_b200.h).  Only
tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import
this module; the product (dde_b200/) never does.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
REF_SO = HERE / "_ref" / "libdde_ref.so"
ORACLE_SO = HERE / "liboracle.so"

c_double_p = C.POINTER(C.c_double)
c_int32_p = C.POINTER(C.c_int32)
c_uint64_p = C.POINTER(C.c_uint64)
_SZ = C.c_size_t


class Control(C.Structure):
    # struct dde_dopri_control, include/dde_b200.h
    _fields_ = [
        ("atol", C.c_double), ("rtol", C.c_double), ("step_size_min", C.c_double),
        ("step_size_max", C.c_double), ("step_size_initial", C.c_double),
        ("step_max_n", C.c_uint64), ("stiff_check", C.c_uint64), ("n_history", C.c_uint64),
        ("method", C.c_int32), ("step_size_min_allow", C.c_int32), ("grow_history", C.c_int32),
        ("return_initial", C.c_int32),
    ]


def control(method="dopri5", rtol=1e-6, atol=1e-6, step_size_min=0.0, step_size_max=float("inf"),
            step_size_initial=0.0, step_max_n=100000, step_size_min_allow=False, stiff_check=0,
            n_history=0, grow_history=False, return_initial=True):
    return Control(atol, rtol, step_size_min, step_size_max, step_size_initial, step_max_n,
                   stiff_check, n_history, {"dopri5": 0, "dopri853": 1}[method],
                   int(step_size_min_allow), int(grow_history), int(return_initial))


_DOPRI_ARGS = [C.c_char_p, C.POINTER(Control), _SZ, _SZ, _SZ, c_double_p, c_double_p, _SZ, _SZ,
               c_double_p, _SZ, c_double_p, _SZ, c_double_p, c_double_p, c_int32_p, c_int32_p,
               c_double_p, c_double_p]
_DIFEQ_ARGS = [C.c_char_p, _SZ, _SZ, _SZ, c_double_p, c_double_p, _SZ, _SZ, c_uint64_p, _SZ, _SZ,
               C.c_int32, c_double_p, c_double_p, c_int32_p, c_double_p]

_libs = {}


def available(backend):
    return (REF_SO if backend == "ref" else ORACLE_SO).exists()


def load(backend):
    if backend in _libs:
        return _libs[backend]
    path = REF_SO if backend == "ref" else ORACLE_SO
    lib = C.CDLL(str(path), mode=C.RTLD_GLOBAL if backend == "ref" else C.RTLD_LOCAL)
    pre = "ref" if backend == "ref" else "oracle"
    for name, args in (("dopri_batch", _DOPRI_ARGS), ("dopri_batch_mp", _DOPRI_ARGS + [C.c_int]),
                       ("difeq_batch", _DIFEQ_ARGS), ("difeq_batch_ex", _DIFEQ_ARGS + [C.c_int])):
        if hasattr(lib, pre + "_" + name):
            fn = getattr(lib, pre + "_" + name)
            fn.restype = C.c_int
            fn.argtypes = args
    fn = getattr(lib, pre + "_dopri_history")
    fn.restype = C.c_long
    fn.argtypes = [C.c_char_p, C.POINTER(Control), _SZ, _SZ, c_double_p, c_double_p, c_double_p,
                   _SZ, c_double_p, _SZ, c_double_p, c_double_p, c_int32_p, c_int32_p, c_double_p,
                   c_double_p, c_double_p, _SZ]
    getattr(lib, pre + "_last_seconds").restype = C.c_double
    getattr(lib, pre + "_last_message").restype = C.c_char_p
    _libs[backend] = lib
    return lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


class Result:
    pass


def _prep(y, parms):
    y = np.ascontiguousarray(y, dtype=np.float64)
    if y.ndim == 1:
        y = y[None, :]
    n_traj, n = y.shape
    if parms is None:
        p = np.zeros((1,), dtype=np.float64)
        n_parms, stride = 0, 0
    else:
        p = np.ascontiguousarray(parms, dtype=np.float64)
        if p.ndim == 1:
            n_parms, stride = p.shape[0], 0
        else:
            n_parms, stride = p.shape[1], p.shape[1]
    return y, p, n_traj, n, n_parms, stride


def set_events(backend, events):
    """Events for the dopri calls that follow: one model event index (or -1) per
    critical time; None clears."""
    lib = load(backend)
    fn = getattr(lib, ("ref" if backend == "ref" else "oracle") + "_set_events")
    fn.restype = None
    fn.argtypes = [c_int32_p, _SZ]
    if events is None:
        fn(None, 0)
    else:
        ev = np.ascontiguousarray(events, dtype=np.int32)
        fn(ev.ctypes.data_as(c_int32_p), ev.shape[0])


def dopri_batch(backend, model, y, times, parms=None, *, n_out=0, tcrit=None, n_proc=0,
                want_y=True, events=None, **ctl_kwargs):
    """Run every row of y through the checker, serially (n_proc = 0) or over
    n_proc forked processes (backend "ref" only)."""
    lib = load(backend)
    pre = "ref" if backend == "ref" else "oracle"
    ctl = control(**ctl_kwargs)
    y, p, n_traj, n, n_parms, stride = _prep(y, parms)
    times = np.ascontiguousarray(times, dtype=np.float64)
    tc = None if tcrit is None else np.ascontiguousarray(tcrit, dtype=np.float64)
    nt = times.shape[0] if ctl.return_initial else times.shape[0] - 1
    r = Result()
    r.y = np.zeros((n_traj, nt, n)) if want_y else None
    r.output = np.zeros((n_traj, nt, n_out)) if (n_out > 0 and want_y) else None
    r.statistics = np.zeros((n_traj, 4), dtype=np.int32)
    r.code = np.zeros((n_traj,), dtype=np.int32)
    r.t_last = np.zeros((n_traj,))
    r.step_size = np.zeros((n_traj,))
    args = [model.encode(), C.byref(ctl), n, n_out, n_traj, _dp(y), _dp(p), n_parms, stride,
            _dp(times), times.shape[0], _dp(tc), 0 if tc is None else tc.shape[0], _dp(r.y),
            _dp(r.output), _ip(r.statistics), _ip(r.code), _dp(r.t_last), _dp(r.step_size)]
    if events is not None:
        p = p.copy()  # event functions may change the parameters in place
        args[6] = _dp(p)
        set_events(backend, events)
    try:
        if n_proc > 0:
            rc = getattr(lib, pre + "_dopri_batch_mp")(*args, n_proc)
        else:
            rc = getattr(lib, pre + "_dopri_batch")(*args)
    finally:
        if events is not None:
            set_events(backend, None)
    if rc != 0:
        raise RuntimeError(getattr(lib, pre + "_last_message")().decode())
    r.seconds = getattr(lib, pre + "_last_seconds")()
    return r


def dopri_continue(backend, model, y, times1, times2, parms=None, *, y_new=None, n_out=0,
                   tcrit2=None, **ctl_kwargs):
    """dopri() over times1, then dopri_continue() over times2 on the same solver
    object (src/r_dopri.c:193-262); the results are those of the second segment."""
    lib = load(backend)
    pre = "ref" if backend == "ref" else "oracle"
    fn = getattr(lib, pre + "_dopri_batch_continue")
    fn.restype = C.c_int
    fn.argtypes = [C.c_char_p, C.POINTER(Control), _SZ, _SZ, _SZ, c_double_p, c_double_p, _SZ,
                   _SZ, c_double_p, _SZ, c_double_p, _SZ, c_double_p, c_double_p, _SZ, c_double_p,
                   c_double_p, c_int32_p, c_int32_p, c_double_p, c_double_p]
    ctl = control(**ctl_kwargs)
    y, p, n_traj, n, n_parms, stride = _prep(y, parms)
    t1 = np.ascontiguousarray(times1, dtype=np.float64)
    t2 = np.ascontiguousarray(times2, dtype=np.float64)
    tc = None if tcrit2 is None else np.ascontiguousarray(tcrit2, dtype=np.float64)
    yn = None if y_new is None else np.ascontiguousarray(y_new, dtype=np.float64).reshape(n_traj, n)
    nt = t2.shape[0] if ctl.return_initial else t2.shape[0] - 1
    r = Result()
    r.y = np.zeros((n_traj, nt, n))
    r.output = np.zeros((n_traj, nt, n_out)) if n_out > 0 else None
    r.statistics = np.zeros((n_traj, 4), dtype=np.int32)
    r.code = np.zeros((n_traj,), dtype=np.int32)
    r.t_last = np.zeros((n_traj,))
    r.step_size = np.zeros((n_traj,))
    rc = fn(model.encode(), C.byref(ctl), n, n_out, n_traj, _dp(y), _dp(p), n_parms, stride,
            _dp(t1), t1.shape[0], _dp(t2), t2.shape[0], _dp(yn), _dp(tc),
            0 if tc is None else tc.shape[0], _dp(r.y), _dp(r.output), _ip(r.statistics),
            _ip(r.code), _dp(r.t_last), _dp(r.step_size))
    if rc != 0:
        raise RuntimeError(getattr(lib, pre + "_last_message")().decode())
    return r


def difeq_history(backend, model, y, steps, parms=None, *, n_out=0, n_history=0,
                  grow_history=False, return_initial=True, max_entries=None):
    """One replicate, also returning the `history` attribute [n_entries, 1 + n + n_out]."""
    lib = load(backend)
    pre = "ref" if backend == "ref" else "oracle"
    y, p, n_rep, n, n_parms, stride = _prep(y, parms)
    assert n_rep == 1
    steps = np.ascontiguousarray(steps, dtype=np.uint64)
    nt = steps.shape[0] if return_initial else steps.shape[0] - 1
    hl = 1 + n + n_out
    cap = max(int(max_entries or n_history), 1)
    r = Result()
    r.y = np.zeros((nt, n))
    r.output = np.zeros((nt, n_out)) if n_out > 0 else None
    r.code = np.zeros((1,), dtype=np.int32)
    hist = np.zeros((cap, hl))
    fn = getattr(lib, pre + "_difeq_history")
    fn.restype = C.c_long
    used = fn(model.encode(), C.c_size_t(n), C.c_size_t(n_out), _dp(y), _dp(p),
              steps.ctypes.data_as(c_uint64_p), C.c_size_t(steps.shape[0]),
              C.c_size_t(n_history), C.c_int32(int(grow_history)), C.c_int32(int(return_initial)),
              _dp(r.y), _dp(r.output), _ip(r.code), _dp(hist), C.c_size_t(cap))
    if used < 0:
        raise RuntimeError("difeq_history failed")
    r.n_entries = int(used)
    r.history = hist[:min(used, cap)].copy()
    return r


def dopri_history(backend, model, y, times, parms=None, *, n_out=0, tcrit=None, max_entries=None,
                  **ctl_kwargs):
    """One trajectory, also returning the `history` attribute [n_entries, order*n + 2]."""
    lib = load(backend)
    pre = "ref" if backend == "ref" else "oracle"
    ctl = control(**ctl_kwargs)
    y, p, n_traj, n, n_parms, stride = _prep(y, parms)
    assert n_traj == 1
    times = np.ascontiguousarray(times, dtype=np.float64)
    tc = None if tcrit is None else np.ascontiguousarray(tcrit, dtype=np.float64)
    nt = times.shape[0] if ctl.return_initial else times.shape[0] - 1
    order = 5 if ctl.method == 0 else 8
    hl = order * n + 2
    cap = max(int(max_entries or ctl.n_history), 1)
    r = Result()
    r.y = np.zeros((nt, n))
    r.output = np.zeros((nt, n_out)) if n_out > 0 else None
    r.statistics = np.zeros((4,), dtype=np.int32)
    r.code = np.zeros((1,), dtype=np.int32)
    r.t_last = np.zeros((1,))
    r.step_size = np.zeros((1,))
    hist = np.zeros((cap, hl))
    used = getattr(lib, pre + "_dopri_history")(
        model.encode(), C.byref(ctl), n, n_out, _dp(y), _dp(p), _dp(times), times.shape[0],
        _dp(tc), 0 if tc is None else tc.shape[0], _dp(r.y), _dp(r.output), _ip(r.statistics),
        _ip(r.code), _dp(r.t_last), _dp(r.step_size), _dp(hist), cap)
    if used < 0:
        raise RuntimeError(getattr(lib, pre + "_last_message")().decode())
    r.n_entries = int(used)
    r.history = hist[:min(used, cap)].copy()
    return r


def difeq_continue(backend, model, y, steps1, steps2, parms=None, *, y_new=None, n_out=0,
                   n_history=0, return_initial=True):
    """difeq() over steps1 then difeq_continue() over steps2 on the same object
    (src/r_difeq.c:114-165); results of the second segment."""
    lib = load(backend)
    pre = "ref" if backend == "ref" else "oracle"
    fn = getattr(lib, pre + "_difeq_batch_continue")
    fn.restype = C.c_int
    fn.argtypes = [C.c_char_p, _SZ, _SZ, _SZ, c_double_p, c_double_p, _SZ, _SZ, c_uint64_p, _SZ,
                   c_uint64_p, _SZ, c_double_p, _SZ, C.c_int32, c_double_p, c_double_p, c_int32_p,
                   c_double_p]
    y, p, n_rep, n, n_parms, stride = _prep(y, parms)
    s1 = np.ascontiguousarray(steps1, dtype=np.uint64)
    s2 = np.ascontiguousarray(steps2, dtype=np.uint64)
    yn = None if y_new is None else np.ascontiguousarray(y_new, dtype=np.float64).reshape(n_rep, n)
    nt = s2.shape[0] if return_initial else s2.shape[0] - 1
    r = Result()
    r.y = np.zeros((n_rep, nt, n))
    r.output = np.zeros((n_rep, nt, n_out)) if n_out > 0 else None
    r.code = np.zeros((n_rep,), dtype=np.int32)
    r.y_final = np.zeros((n_rep, n))
    rc = fn(model.encode(), n, n_out, n_rep, _dp(y), _dp(p), n_parms, stride,
            s1.ctypes.data_as(c_uint64_p), s1.shape[0], s2.ctypes.data_as(c_uint64_p), s2.shape[0],
            _dp(yn), n_history, int(return_initial), _dp(r.y), _dp(r.output), _ip(r.code),
            _dp(r.y_final))
    if rc != 0:
        raise RuntimeError(getattr(lib, pre + "_last_message")().decode())
    return r


def difeq_batch(backend, model, y, steps, parms=None, *, n_out=0, n_history=0,
                return_initial=True, shared_ring=False):
    lib = load(backend)
    pre = "ref" if backend == "ref" else "oracle"
    y, p, n_rep, n, n_parms, stride = _prep(y, parms)
    steps = np.ascontiguousarray(steps, dtype=np.uint64)
    nt = steps.shape[0] if return_initial else steps.shape[0] - 1
    r = Result()
    r.y = np.zeros((n_rep, nt, n))
    r.output = np.zeros((n_rep, nt, n_out)) if n_out > 0 else None
    r.code = np.zeros((n_rep,), dtype=np.int32)
    r.y_final = np.zeros((n_rep, n))
    args = [model.encode(), n, n_out, n_rep, _dp(y), _dp(p), n_parms, stride,
            steps.ctypes.data_as(c_uint64_p), steps.shape[0], n_history, int(return_initial),
            _dp(r.y), _dp(r.output), _ip(r.code), _dp(r.y_final)]
    if shared_ring:
        rc = getattr(lib, pre + "_difeq_batch_ex")(*args, 1)
    else:
        rc = getattr(lib, pre + "_difeq_batch")(*args)
    if rc != 0:
        raise RuntimeError(getattr(lib, pre + "_last_message")().decode())
    r.seconds = getattr(lib, pre + "_last_seconds")()
    return r
