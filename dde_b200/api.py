"""Host-side mirror of the reference's R interface, over the C-ABI.

R is not available in this image, so where the reference has R wrappers
around `.Call(Cdopri, ...)` / `.Call(Cdifeq, ...)` this module has Python
functions with the same names, argument meaning and error behaviour
(/root/reference/R/dopri.R:293-437, R/difeq.R:71-141,
R/difeq_replicate.R:47-152) calling the batched C-ABI (include/dde_b200.h).
`func` / `target` name a compiled model (the reference's "name of a compiled
function" + `dllname` route, R/common.R:89-98); R-closure models have no
device equivalent.

Single-trajectory calls (`dopri`, `difeq`) return what the R functions
return (a matrix with a leading time/step column by default) as an ndarray
subclass carrying the attributes (`statistics`, `step_size`, `history`,
`output`), and raise the reference's error messages.  The batched calls
(`dopri_batch`, `difeq_replicate`) never raise for a per-trajectory failure:
codes come back per trajectory.
"""
import ctypes as C
import math

import numpy as np

from . import _lib
from ._lib import DopriControl, c_double_p, c_int32_p, c_uint64_p

DOPRI_5 = 0
DOPRI_853 = 1

ERR_ZERO_TIME_DIFFERENCE = -1
ERR_INCONSISTENT_TIME = -2
ERR_TOO_MANY_STEPS = -3
ERR_STEP_SIZE_TOO_SMALL = -4
ERR_STEP_SIZE_VANISHED = -5
ERR_YLAG_FAIL = -6
ERR_STIFF = -7
ERR_NONFINITE_DERIV = -8
ERR_DIFEQ_HISTORY = -9
OK_COMPLETE = 1

STAT_NAMES = ("n_eval", "n_step", "n_accept", "n_reject")


class DdeError(RuntimeError):
    """An error the reference would raise with Rf_error / stop()."""


def _dp(a):
    return None if a is None else a.ctypes.data_as(c_double_p)


def _ip(a):
    return None if a is None else a.ctypes.data_as(c_int32_p)


def _f64(a, name):
    arr = np.ascontiguousarray(a, dtype=np.float64)
    if not np.issubdtype(np.asarray(a).dtype, np.number) and np.asarray(a).dtype != bool:
        raise DdeError("'%s' must be numeric" % name)
    return arr


def _method_code(method):
    # match_value(method, dopri_methods()), R/dopri.R:341 and R/dopri.R:710-712
    methods = {"dopri5": DOPRI_5, "dopri853": DOPRI_853}
    if method not in methods:
        raise DdeError("'method' must be one of {%s}" % ", ".join(methods))
    return methods[method]


def _control(method, rtol, atol, step_size_min, step_size_max, step_size_initial, step_max_n,
             step_size_min_allow, stiff_check, n_history, grow_history, return_initial):
    for name, val in (("rtol", rtol), ("atol", atol), ("step_size_min", step_size_min),
                      ("step_size_max", step_size_max), ("step_size_initial", step_size_initial)):
        if np.ndim(val) != 0:
            raise DdeError("'%s' must be a scalar" % name)  # assert_scalar, R/util.R
    for name, val in (("step_max_n", step_max_n), ("n_history", n_history),
                      ("stiff_check", stiff_check)):
        if np.ndim(val) != 0 or val < 0 or int(val) != val:
            raise DdeError("'%s' must be a non-negative integer" % name)  # assert_size
    return DopriControl(float(atol), float(rtol), float(step_size_min), float(step_size_max),
                        float(step_size_initial), int(step_max_n), int(stiff_check),
                        int(n_history), _method_code(method), int(bool(step_size_min_allow)),
                        int(bool(grow_history)), int(bool(return_initial)))


def strerror(code, t=0.0, n_history=0):
    """The message the reference raises for a return code (src/r_dopri.c:264-297)."""
    lib = _lib.load()
    buf = C.create_string_buffer(256)
    lib.dde_dopri_strerror(int(code), float(t), int(n_history), buf, 256)
    return buf.value.decode()


def models():
    """Names of the compiled models known to the library."""
    lib = _lib.load()
    return sorted({lib.dde_model_name(i).decode() for i in range(lib.dde_model_count())})


class DopriBatchResult:
    """Results of a batched solve, in the reference's array layouts with the
    trajectory as the slowest dimension."""

    def __init__(self, y, output, statistics, code, t_last, step_size, times, n_history):
        self.y = y                    # [n_traj, nt, n]   (R: n x nt x n_traj)
        self.output = output          # [n_traj, nt, n_out] or None
        self.statistics = statistics  # [n_traj, 4]: n_eval, n_step, n_accept, n_reject
        self.code = code              # [n_traj] return_code, src/dopri.h:24-35
        self.t_last = t_last          # [n_traj]
        self.step_size = step_size    # [n_traj]
        self.times = times
        self.n_history = n_history

    @property
    def ok(self):
        return self.code == OK_COMPLETE

    def message(self, j):
        return strerror(self.code[j], self.t_last[j], self.n_history)


def _device_list(devices):
    """"all" -> None (every visible device); else an int32 array of indices."""
    if isinstance(devices, str):
        if devices != "all":
            raise DdeError("'devices' must be None, \"all\" or a list of device indices")
        return None
    dev = np.ascontiguousarray(np.atleast_1d(devices), dtype=np.int32)
    if dev.shape[0] == 0:
        raise DdeError("'devices' must not be empty")
    return dev


def _prep_batch_inputs(y, parms, n_traj_hint=None):
    y = _f64(y, "y")
    if y.ndim == 1:
        y = y[None, :]
    if y.ndim != 2:
        raise DdeError("'y' must be [n_traj, n]")
    n_traj, n = y.shape
    if parms is None:
        parms_arr = np.zeros((0,), dtype=np.float64)
        n_parms, stride = 0, 0
    else:
        parms_arr = _f64(parms, "parms")
        if parms_arr.ndim == 0:
            parms_arr = parms_arr[None]
        if parms_arr.ndim == 1:
            n_parms, stride = parms_arr.shape[0], 0  # one vector shared by all trajectories
        elif parms_arr.ndim == 2 and parms_arr.shape[0] == n_traj:
            n_parms, stride = parms_arr.shape[1], parms_arr.shape[1]
        else:
            raise DdeError("'parms' must be [n_parms] or [n_traj, n_parms]")
    return y, parms_arr, n_traj, n, n_parms, stride


def dopri_batch(y, times, func, parms=None, *, n_out=0, rtol=1e-6, atol=1e-6, step_size_min=0.0,
                step_size_max=math.inf, step_size_initial=0.0, step_max_n=100000,
                step_size_min_allow=False, tcrit=None, method="dopri5", stiff_check=0,
                n_history=0, grow_history=False, return_initial=True, devices=None,
                events=None):
    """Integrate n_traj independent trajectories of one compiled model.

    One call == the reference's `dopri()` run once per row of `y` (and per row
    of `parms`), concurrently on the GPU.  Arguments as R/dopri.R:293-310.
    `devices`: None = the current device; "all" or a list of device indices =
    shard the trajectories over those GPUs (dde_dopri_batch_sharded).
    `events`: one entry per critical time -- the position of the event function
    in the model's event table, or -1 for a plain stop (R/events.R merges
    `event_time` into `tcrit` the same way).
    """
    lib = _lib.load()
    y, parms_arr, n_traj, n, n_parms, stride = _prep_batch_inputs(y, parms)
    times = _f64(times, "times")
    if times.ndim != 1 or times.shape[0] < 2:
        raise DdeError("At least two times must be given")  # src/r_dopri.c:51-53
    tc = None if tcrit is None else _f64(tcrit, "tcrit").ravel()
    ctl = _control(method, rtol, atol, step_size_min, step_size_max, step_size_initial, step_max_n,
                   step_size_min_allow, stiff_check, n_history, grow_history, return_initial)
    nt = times.shape[0] if return_initial else times.shape[0] - 1
    y_out = np.empty((n_traj, nt, n), dtype=np.float64)
    out = np.empty((n_traj, nt, n_out), dtype=np.float64) if n_out > 0 else None
    stats = np.empty((n_traj, 4), dtype=np.int32)
    code = np.empty((n_traj,), dtype=np.int32)
    t_last = np.empty((n_traj,), dtype=np.float64)
    step_size = np.empty((n_traj,), dtype=np.float64)
    args = [func.encode(), C.byref(ctl), n, n_out, n_traj, _dp(y),
            _dp(parms_arr) if n_parms else None, n_parms, stride, _dp(times), times.shape[0],
            _dp(tc), 0 if tc is None else tc.shape[0], _dp(y_out), _dp(out), _ip(stats),
            _ip(code), _dp(t_last), _dp(step_size)]
    if events is not None and devices is not None:
        ev = np.ascontiguousarray(events, dtype=np.int32).ravel()
        if tc is None or ev.shape[0] != tc.shape[0]:
            raise DdeError("one event index per critical time is required")
        dev = _device_list(devices)
        rc = lib.dde_dopri_batch_sharded_events(*args[:13], _ip(ev), *args[13:], _ip(dev),
                                                0 if dev is None else dev.shape[0])
        if rc != 0:
            raise DdeError(_lib.last_error())
        return DopriBatchResult(y_out, out, stats, code, t_last, step_size, times, int(n_history))
    if events is not None:
        from .batch import DopriBatch
        b = DopriBatch(func, n, n_traj, n_parms=n_parms, n_out=n_out, method=method, rtol=rtol,
                       atol=atol, step_size_min=step_size_min, step_size_max=step_size_max,
                       step_size_initial=step_size_initial, step_max_n=step_max_n,
                       step_size_min_allow=step_size_min_allow, stiff_check=stiff_check,
                       n_history=n_history, return_initial=return_initial)
        try:
            b.upload(y, parms_arr if n_parms else None, shared_parms=(stride == 0))
            b.set_times(times, tc)
            b.set_events(events)
            b.run()
            return b.download()
        finally:
            b.close()
    if devices is None:
        rc = lib.dde_dopri_batch(*args)
    else:
        dev = _device_list(devices)
        rc = lib.dde_dopri_batch_sharded(*args, _ip(dev), 0 if dev is None else dev.shape[0])
    if rc != 0:
        raise DdeError(_lib.last_error())
    return DopriBatchResult(y_out, out, stats, code, t_last, step_size, times, int(n_history))


class DopriResult(np.ndarray):
    """ndarray with the attributes the R function attaches (src/r_dopri.c:394-441)."""

    statistics = None
    step_size = None
    history = None
    output = None
    ptr = None

    def __array_finalize__(self, obj):
        if obj is None:
            return
        for name in ("statistics", "step_size", "history", "output", "ptr"):
            setattr(self, name, getattr(obj, name, None))


class DopriHistory(np.ndarray):
    """The `dopri_history` matrix: (order*n + 2) x n_entries, attribute `n`
    (src/r_dopri.c:401-410); stored row = entry here."""

    n = None

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.n = getattr(obj, "n", None)


def _bind(times, y, output, return_time, return_output_with_y, return_by_column, return_initial):
    # prepare_output, R/common.R:36-86; y is [nt, n] here (already "by column")
    cols = []
    if return_time:
        cols.append((times if return_initial else times[1:])[:, None])
    cols.append(y)
    bind_output = output is not None and return_output_with_y
    if bind_output:
        cols.append(output)
    ret = np.concatenate(cols, axis=1) if len(cols) > 1 else y
    if not return_by_column:
        ret = ret.T
    return ret, (None if bind_output else output)


def dopri(y, times, func, parms=None, *, n_out=0, rtol=1e-6, atol=1e-6, step_size_min=0.0,
          step_size_max=math.inf, step_size_initial=0.0, step_max_n=100000,
          step_size_min_allow=False, tcrit=None, method="dopri5", stiff_check=0, n_history=0,
          grow_history=False, return_history=None, return_by_column=True, return_initial=True,
          return_time=True, return_output_with_y=True, return_statistics=False,
          return_minimal=False, restartable=False, _resume=None):
    """One trajectory; R/dopri.R:293-437.  Raises DdeError with the reference's
    message on failure (src/r_dopri.c:264-297).  `restartable=True` keeps the
    solver object alive in the result's `ptr` attribute for `dopri_continue`."""
    if return_history is None:
        return_history = n_history > 0
    if return_minimal:  # R/dopri.R:317-323
        return_by_column = False
        return_initial = False
        return_time = False
        return_output_with_y = False
    y = _f64(y, "y").ravel()
    times_arr = _f64(times, "times").ravel()
    if times_arr.shape[0] < 2:
        raise DdeError("At least two times must be given")
    lib = _lib.load()
    ctl = _control(method, rtol, atol, step_size_min, step_size_max, step_size_initial, step_max_n,
                   step_size_min_allow, stiff_check, n_history, grow_history, return_initial)
    n = y.shape[0]
    parms_arr = None if parms is None else np.atleast_1d(_f64(parms, "parms")).ravel()
    n_parms = 0 if parms_arr is None else parms_arr.shape[0]
    tc = None if tcrit is None else _f64(tcrit, "tcrit").ravel()
    keep_state = False
    if _resume is None:
        h = lib.dde_dopri_batch_alloc(func.encode(), C.byref(ctl), n, n_out, n_parms, 1)
        if not h:
            raise DdeError(_lib.last_error())
    else:
        h, keep_state = _resume
    keep = False
    try:
        nt = times_arr.shape[0] if return_initial else times_arr.shape[0] - 1
        y_out = np.empty((nt, n))
        out = np.empty((nt, n_out)) if n_out > 0 else None
        stats = np.empty((4,), dtype=np.int32)
        code = np.empty((1,), dtype=np.int32)
        t_last = np.empty((1,))
        step_size = np.empty((1,))
        if _resume is None:
            rc = lib.dde_dopri_batch_upload(h, y.ctypes.data, None if parms_arr is None else
                                            parms_arr.ctypes.data, 0)
            rc = rc or lib.dde_dopri_batch_set_times(h, _dp(times_arr), times_arr.shape[0],
                                                     _dp(tc), 0 if tc is None else tc.shape[0])
            rc = rc or lib.dde_dopri_batch_run(h)
        else:  # dopri_continue: y is None = carry on from the final state
            rc = lib.dde_dopri_batch_continue(h, None if keep_state else y.ctypes.data,
                                              _dp(times_arr), times_arr.shape[0], _dp(tc),
                                              0 if tc is None else tc.shape[0])
        rc = rc or lib.dde_dopri_batch_download(h, y_out.ctypes.data, None if out is None else
                                                out.ctypes.data, stats.ctypes.data,
                                                code.ctypes.data, t_last.ctypes.data,
                                                step_size.ctypes.data)
        if rc != 0:
            raise DdeError(_lib.last_error())
        if code[0] != OK_COMPLETE:
            raise DdeError(strerror(code[0], t_last[0], n_history))
        history = None
        if return_history:
            order = 5 if ctl.method == DOPRI_5 else 8
            hl = order * n + 2
            cap = max(int(lib.dde_dopri_batch_n_history(h)), 1)  # grow_history may enlarge it
            buf = np.empty((cap, hl))
            used = lib.dde_dopri_batch_history(h, 0, _dp(buf), cap)
            if used < 0:
                raise DdeError(_lib.last_error())
            history = buf[:used].copy().view(DopriHistory)
            history.n = n
        keep = bool(restartable)
    finally:
        if not keep:
            lib.dde_dopri_batch_free(h)
    ret, output_attr = _bind(times_arr, y_out, out, return_time, return_output_with_y,
                             return_by_column, return_initial)
    ret = np.ascontiguousarray(ret).view(DopriResult)
    ret.output = output_attr if (output_attr is None or return_by_column) else output_attr.T
    ret.history = history
    if return_statistics:
        ret.statistics = dict(zip(STAT_NAMES, (int(v) for v in stats)))
        ret.step_size = float(step_size[0])
    if keep:  # the "ptr" / "restart_data" attributes, R/dopri.R:424-436
        ret.ptr = _DopriPtr(h, dict(func=func, n=n, n_out=n_out, rtol=rtol, atol=atol,
                                    step_size_min=step_size_min, step_size_max=step_size_max,
                                    step_max_n=step_max_n, step_size_min_allow=step_size_min_allow,
                                    method=method, stiff_check=stiff_check, n_history=n_history,
                                    t_last=float(t_last[0])))
    return ret


class _DopriPtr:
    """The live solver object behind a restartable result (R: an external
    pointer with a finaliser, src/r_dopri.c:375-392)."""

    def __init__(self, handle, meta):
        self.handle = handle
        self.meta = meta

    def __del__(self):
        try:
            if self.handle:
                _lib.load().dde_dopri_batch_free(self.handle)
                self.handle = None
        except Exception:
            pass


def dopri_continue(obj, times, y=None, *, copy=False, tcrit=None, return_history=None,
                   return_by_column=True, return_initial=True, return_time=True,
                   return_output_with_y=True, return_statistics=False, restartable=False):
    """Continue a restartable `dopri()` result; R/dopri.R:469-511."""
    ptr = getattr(obj, "ptr", None)
    if not isinstance(ptr, _DopriPtr) or not ptr.handle:
        raise DdeError("Expected an object created with restartable = TRUE")
    lib = _lib.load()
    m = ptr.meta
    handle = ptr.handle
    if copy:  # dopri_copy, then continue the copy: the original stays usable
        handle = lib.dde_dopri_batch_copy(ptr.handle)
        if not handle:
            raise DdeError(_lib.last_error())
    elif restartable:
        ptr.handle = None  # ownership moves to the new result
    times_arr = _f64(times, "times").ravel()
    if times_arr.shape[0] >= 1 and times_arr[0] != m["t_last"]:
        if copy:
            lib.dde_dopri_batch_free(handle)
        elif restartable:
            ptr.handle = handle
        raise DdeError("Incorrect initial time on integration restart")  # src/r_dopri.c:216
    if y is not None and np.size(y) != m["n"]:
        raise DdeError("Incorrect size 'y' on integration restart")  # src/r_dopri.c:203
    keep_alive = restartable or not copy
    ret = dopri(np.zeros(m["n"]) if y is None else y, times_arr, m["func"], None, n_out=m["n_out"],
                rtol=m["rtol"], atol=m["atol"], step_size_min=m["step_size_min"],
                step_size_max=m["step_size_max"], step_max_n=m["step_max_n"],
                step_size_min_allow=m["step_size_min_allow"], tcrit=tcrit, method=m["method"],
                stiff_check=m["stiff_check"], n_history=m["n_history"],
                return_history=return_history, return_by_column=return_by_column,
                return_initial=return_initial, return_time=return_time,
                return_output_with_y=return_output_with_y, return_statistics=return_statistics,
                restartable=keep_alive, _resume=(handle, y is None))
    if not restartable:
        if copy:
            ret.ptr = None  # the temporary copy is freed with its _DopriPtr
        else:
            # the original object keeps its (advanced) solver, as in R without copy
            ptr.handle = ret.ptr.handle
            ptr.meta = ret.ptr.meta
            ret.ptr.handle = None
            ret.ptr = None
    return ret


def dopri5(y, times, func, parms=None, **kwargs):
    """R/dopri.R:441-443"""
    return dopri(y, times, func, parms, method="dopri5", **kwargs)


def dopri853(y, times, func, parms=None, **kwargs):
    """R/dopri.R:447-449"""
    return dopri(y, times, func, parms, method="dopri853", **kwargs)


def dopri_interpolate(h, t):
    """Dense output at times `t` from a stored history; R/dopri.R:577-642.

    `h` is a DopriHistory (or a result carrying one).  Returns [len(t), n].
    Accepts a stack of histories [n_traj, n_entries, order*n + 2] with `n`
    given as attribute or inferred for a batched evaluation, returning
    [n_traj, len(t), n].
    """
    lib = _lib.load()
    if isinstance(h, DopriResult):
        h = h.history
    if h is None:
        raise DdeError("Corrupt history object: 'n' is missing")
    n = getattr(h, "n", None)
    if n is None:
        raise DdeError("Corrupt history object: 'n' is missing")
    arr = np.ascontiguousarray(np.asarray(h), dtype=np.float64)
    single = arr.ndim == 2
    if single:
        arr = arr[None]
    n_traj, n_entries, hl = arr.shape
    t = np.atleast_1d(_f64(t, "t")).ravel()
    y = np.empty((n_traj, n, t.shape[0]))
    rc = lib.dde_dopri_interpolate_batch(_dp(arr), hl, n_entries, n, n_traj, _dp(t), t.shape[0],
                                         _dp(y))
    if rc != 0:
        raise DdeError(_lib.last_error())
    y = np.ascontiguousarray(np.swapaxes(y, 1, 2))  # -> [n_traj, n_t, n]
    return y[0] if single else y


# ---- difeq ------------------------------------------------------------------


class DifeqResult(np.ndarray):
    output = None
    code = None
    history = None

    def __array_finalize__(self, obj):
        if obj is None:
            return
        self.output = getattr(obj, "output", None)
        self.code = getattr(obj, "code", None)
        self.history = getattr(obj, "history", None)


def _steps(steps):
    steps = np.atleast_1d(np.asarray(steps))
    if steps.shape[0] >= 1 and steps[0] < 0:
        raise DdeError("steps must be positive")  # R/difeq.R:112-114
    if steps.shape[0] == 1:
        steps = np.arange(0, int(steps[0]) + 1)  # R/difeq.R:115-117
    return np.ascontiguousarray(steps, dtype=np.uint64)


def _difeq_call(y2d, steps, target, parms_arr, n_parms, stride, n_out, n_history, return_initial,
                devices=None):
    lib = _lib.load()
    n_rep, n = y2d.shape
    if n_history > 0 and n_history < 2:
        raise DdeError("If given, n_history must be at least 2")  # R/difeq.R:119-121
    nt = steps.shape[0] if return_initial else steps.shape[0] - 1
    y_out = np.empty((n_rep, nt, n))
    out = np.empty((n_rep, nt, n_out)) if n_out > 0 else None
    code = np.empty((n_rep,), dtype=np.int32)
    y_final = np.empty((n_rep, n))
    args = [target.encode(), n, n_out, n_rep, _dp(y2d), _dp(parms_arr) if n_parms else None,
            n_parms, stride, steps.ctypes.data_as(c_uint64_p), steps.shape[0], int(n_history),
            int(bool(return_initial)), _dp(y_out), _dp(out), _ip(code), _dp(y_final)]
    if devices is None:
        rc = lib.dde_difeq_batch(*args)
    else:
        dev = _device_list(devices)
        rc = lib.dde_difeq_batch_sharded(*args, _ip(dev), 0 if dev is None else dev.shape[0])
    if rc != 0:
        raise DdeError(_lib.last_error())
    return y_out, out, code, y_final


def _grown_history(n_history, grow_history, st):
    """grow_history for a difference equation: a ring that never overwrites is a ring
    with room for every step of the run plus the initial state (src/difeq.c:49-56)."""
    if grow_history and n_history > 0:
        return max(int(n_history), int(st[-1] - st[0]) + 1)
    return int(n_history)


def difeq(y, steps, target, parms=None, *, n_out=0, n_history=0, grow_history=False,
          return_history=None, return_by_column=True, return_initial=True, return_step=True,
          return_output_with_y=True, return_minimal=False):
    """Iterate a difference equation; R/difeq.R:71-141.  `return_history` (default:
    n_history > 0) attaches the `history` attribute, [entries, 1 + n + n_out] = step, y,
    out, oldest first (src/r_difeq.c:178-186; its out columns are NA, see
    dde_difeq_batch_history)."""
    if return_history is None:
        return_history = n_history > 0
    if return_minimal:
        return_by_column = False
        return_initial = False
        return_step = False
        return_output_with_y = False
    y = _f64(y, "y").ravel()
    st = _steps(steps)
    n_history = _grown_history(n_history, grow_history, st)
    y2d, parms_arr, _, _, n_parms, stride = _prep_batch_inputs(y[None, :], parms)
    history = None
    if return_history and n_history > 0:
        from .batch import DifeqBatch
        if n_history < 2:
            raise DdeError("If given, n_history must be at least 2")  # R/difeq.R:119-121
        b = DifeqBatch(target, y.shape[0], 1, n_parms=n_parms, n_out=n_out, n_history=n_history,
                       restartable=True)
        try:
            b.upload(y2d, parms_arr if n_parms else None)
            b.set_steps(st, return_initial)
            b.run()
            y_out, out, code, _ = b.download()
            if code[0] == OK_COMPLETE:
                history = b.history(0)
        finally:
            b.close()
    else:
        y_out, out, code, _ = _difeq_call(y2d, st, target, parms_arr, n_parms, stride, n_out,
                                          n_history, return_initial)
    if code[0] != OK_COMPLETE:
        raise DdeError(strerror(code[0]))
    ret, output_attr = _bind(st.astype(np.float64), y_out[0], None if out is None else out[0],
                             return_step, return_output_with_y, return_by_column, return_initial)
    ret = np.ascontiguousarray(ret).view(DifeqResult)
    ret.output = output_attr if (output_attr is None or return_by_column) else output_attr.T
    ret.history = history
    return ret


def difeq_replicate(n, y, steps, target, parms=None, *, n_out=0, n_history=0, grow_history=False,
                    return_by_column=True, return_initial=True, return_step=True,
                    return_output_with_y=True, return_minimal=False, devices=None):
    """Replicated difference-equation runs; R/difeq_replicate.R:47-152.

    `y` is one initial state (used for all n replicates), or [n, n_eq] with one
    row per replicate (the reference accepts a matrix / list / generator
    function of that shape, R/difeq_replicate.R:67-88).  `parms` likewise may
    be one vector or [n, n_parms].  Returns [nt, cols, n] by column (the R
    array aperm'ed as R/difeq_replicate.R:196-201) with `.code` per replicate.
    """
    if np.ndim(n) != 0 or int(n) != n or n < 1:
        raise DdeError("'n' must be a positive integer")
    n = int(n)
    if return_minimal:
        return_by_column = False
        return_initial = False
        return_step = False
        return_output_with_y = False
    y = _f64(y, "y")
    if y.ndim == 1:
        y2d = np.ascontiguousarray(np.broadcast_to(y, (n, y.shape[0])))
    elif y.ndim == 2 and y.shape[0] == n:
        y2d = y
    else:
        raise DdeError("'y' must have n rows")  # R/difeq_replicate.R:70-72
    st = _steps(steps)
    n_history = _grown_history(n_history, grow_history, st)
    y2d, parms_arr, _, n_eq, n_parms, stride = _prep_batch_inputs(y2d, parms)
    y_out, out, code, _ = _difeq_call(y2d, st, target, parms_arr, n_parms, stride, n_out,
                                      n_history, return_initial, devices)
    blocks = []
    if return_step:
        s = (st if return_initial else st[1:]).astype(np.float64)
        blocks.append(np.broadcast_to(s[None, :, None], (n, s.shape[0], 1)))
    blocks.append(y_out)
    bind_output = out is not None and return_output_with_y
    if bind_output:
        blocks.append(out)
    ret = np.concatenate(blocks, axis=2) if len(blocks) > 1 else y_out  # [n, nt, cols]
    ret = np.transpose(ret, (1, 2, 0)) if return_by_column else np.transpose(ret, (2, 1, 0))
    ret = np.ascontiguousarray(ret).view(DifeqResult)
    ret.code = code
    if out is not None and not bind_output:
        ret.output = np.transpose(out, (1, 2, 0)) if return_by_column else np.transpose(
            out, (2, 1, 0))
    return ret
