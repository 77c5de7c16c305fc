"""Device-resident batch handles (the analogue of n_traj `dopri_data` /
`difeq_data` objects kept alive between calls, src/dopri.h:55-161).

Inputs are uploaded once; `run()` integrates every trajectory from the
resident state, asynchronously on the handle's own CUDA stream; results are
fetched on demand.  Used by bench.py and by the multi-GPU sharding helpers.
"""
import ctypes as C

import numpy as np

from . import _lib
from .api import (DdeError, DopriBatchResult, _control, _f64, _prep_batch_inputs, _steps)
from ._lib import c_double_p, c_uint64_p


def _ptr(a):
    """Address of a numpy array or torch tensor (host memory)."""
    if a is None:
        return None
    if hasattr(a, "data_ptr"):
        return a.data_ptr()
    return a.ctypes.data


class DopriBatch:
    def __init__(self, func, n, n_traj, *, n_parms=0, n_out=0, method="dopri5", rtol=1e-6,
                 atol=1e-6, step_size_min=0.0, step_size_max=float("inf"), step_size_initial=0.0,
                 step_max_n=100000, step_size_min_allow=False, stiff_check=0, n_history=0,
                 grow_history=False, return_initial=True):
        self._lib = _lib.load()
        self.ctl = _control(method, rtol, atol, step_size_min, step_size_max, step_size_initial,
                            step_max_n, step_size_min_allow, stiff_check, n_history, grow_history,
                            return_initial)
        self.n, self.n_traj, self.n_parms, self.n_out = int(n), int(n_traj), int(n_parms), int(n_out)
        self.n_history = int(n_history)
        self.return_initial = bool(return_initial)
        self.times = None
        self._h = self._lib.dde_dopri_batch_alloc(func.encode(), C.byref(self.ctl), self.n,
                                                  self.n_out, self.n_parms, self.n_traj)
        if not self._h:
            raise DdeError(_lib.last_error())

    def _check(self, rc):
        if rc != 0:
            raise DdeError(_lib.last_error())

    def upload(self, y0=None, parms=None, shared_parms=False):
        """Host -> device.  Arrays may be numpy arrays or (pinned) torch CPU
        tensors of float64 in the C-ABI layout: y0 [n_traj, n], parms
        [n_traj, n_parms] (or [n_parms] with shared_parms)."""
        stride = 0 if shared_parms else self.n_parms
        self._check(self._lib.dde_dopri_batch_upload(self._h, _ptr(y0), _ptr(parms), stride))

    def set_times(self, times, tcrit=None):
        times = _f64(times, "times").ravel()
        tc = None if tcrit is None else _f64(tcrit, "tcrit").ravel()
        self._check(self._lib.dde_dopri_batch_set_times(
            self._h, times.ctypes.data_as(c_double_p), times.shape[0],
            None if tc is None else tc.ctypes.data_as(c_double_p), 0 if tc is None else tc.shape[0]))
        self.times = times
        self.nt = times.shape[0] if self.return_initial else times.shape[0] - 1

    def set_events(self, events):
        """One entry per critical time: the model's event function to apply
        there (its position in the model's event table), or -1 for a plain stop."""
        ev = np.ascontiguousarray(events, dtype=np.int32).ravel()
        self._check(self._lib.dde_dopri_batch_set_events(
            self._h, ev.ctypes.data_as(_lib.c_int32_p), ev.shape[0]))

    def set_order(self, order):
        """The order in which trajectories are handed out to the device's threads (a
        permutation of 0 .. n_traj - 1; None: index order).  Results do not depend on it."""
        if order is None:
            self._check(self._lib.dde_dopri_batch_set_order(self._h, None))
            return
        o = np.ascontiguousarray(order, dtype=np.uint32).ravel()
        self._check(self._lib.dde_dopri_batch_set_order(self._h, o.ctypes.data_as(C.c_void_p)))

    def order_by_last_run(self):
        """Longest first, by the step attempts of this handle's last run."""
        self._check(self._lib.dde_dopri_batch_order_by_last_run(self._h))

    def run(self):
        self._check(self._lib.dde_dopri_batch_run(self._h))

    def continue_(self, times, y=None, tcrit=None):
        """dopri_continue (R/dopri.R:469-511): integrate on from where the last
        run ended; `y` [n_traj, n] replaces the final states (and becomes the
        pre-history), times[0] must be the time every trajectory stopped at."""
        times = _f64(times, "times").ravel()
        tc = None if tcrit is None else _f64(tcrit, "tcrit").ravel()
        if y is not None and not hasattr(y, "data_ptr"):
            y = np.ascontiguousarray(y, dtype=np.float64)
            if y.size != self.n * self.n_traj:
                raise DdeError("Incorrect size 'y' on integration restart")  # src/r_dopri.c:203
        self._check(self._lib.dde_dopri_batch_continue(
            self._h, _ptr(y), times.ctypes.data_as(c_double_p), times.shape[0],
            None if tc is None else tc.ctypes.data_as(c_double_p), 0 if tc is None else tc.shape[0]))
        self.times = times
        self.nt = times.shape[0] if self.return_initial else times.shape[0] - 1

    def copy(self):
        """dopri_data_copy (src/dopri.c:92-134): an independent handle with the same state."""
        h = self._lib.dde_dopri_batch_copy(self._h)
        if not h:
            raise DdeError(_lib.last_error())
        other = object.__new__(DopriBatch)
        other.__dict__.update(self.__dict__)
        other._h = h
        return other

    def sync(self):
        self._check(self._lib.dde_dopri_batch_sync(self._h))

    def last_ms(self):
        return self._lib.dde_dopri_batch_last_ms(self._h)

    def last_launches(self):
        return self._lib.dde_dopri_batch_last_launches(self._h)

    def download_into(self, y_out=None, out=None, stats=None, code=None, t_last=None,
                      step_size=None):
        """Device -> host into caller buffers (numpy or pinned torch tensors)."""
        self._check(self._lib.dde_dopri_batch_download(self._h, _ptr(y_out), _ptr(out), _ptr(stats),
                                                       _ptr(code), _ptr(t_last), _ptr(step_size)))

    def download_range(self, first, count, want_y=True):
        """Results of trajectories first .. first + count - 1 only."""
        first, count = int(first), int(count)
        y_out = np.empty((count, self.nt, self.n)) if want_y else None
        out = np.empty((count, self.nt, self.n_out)) if (self.n_out > 0 and want_y) else None
        stats = np.empty((count, 4), dtype=np.int32)
        code = np.empty((count,), dtype=np.int32)
        t_last = np.empty((count,))
        step_size = np.empty((count,))
        self._check(self._lib.dde_dopri_batch_download_range(
            self._h, first, count, _ptr(y_out), _ptr(out), _ptr(stats), _ptr(code), _ptr(t_last),
            _ptr(step_size)))
        return DopriBatchResult(y_out, out, stats, code, t_last, step_size, self.times,
                                self.n_history)

    def run_download_into(self, y_out=None, out=None, stats=None, code=None, t_last=None,
                          step_size=None):
        """run() and download_into() in one blocking call, overlapped: results are copied
        to the host block by block while later trajectories are still being integrated."""
        self._check(self._lib.dde_dopri_batch_run_download(
            self._h, _ptr(y_out), _ptr(out), _ptr(stats), _ptr(code), _ptr(t_last),
            _ptr(step_size)))

    def download(self, want_y=True):
        y_out = np.empty((self.n_traj, self.nt, self.n)) if want_y else None
        out = np.empty((self.n_traj, self.nt, self.n_out)) if (self.n_out > 0 and want_y) else None
        stats = np.empty((self.n_traj, 4), dtype=np.int32)
        code = np.empty((self.n_traj,), dtype=np.int32)
        t_last = np.empty((self.n_traj,))
        step_size = np.empty((self.n_traj,))
        self.download_into(y_out, out, stats, code, t_last, step_size)
        return DopriBatchResult(y_out, out, stats, code, t_last, step_size, self.times,
                                self.n_history)

    def history(self, j):
        order = 5 if self.ctl.method == 0 else 8
        hl = order * self.n + 2
        # the capacity as it is now: grow_history may have enlarged the rings
        self.n_history = int(self._lib.dde_dopri_batch_n_history(self._h))
        cap = max(self.n_history, 1)
        buf = np.empty((cap, hl))
        used = self._lib.dde_dopri_batch_history(self._h, int(j), buf.ctypes.data_as(c_double_p), cap)
        if used < 0:
            raise DdeError(_lib.last_error())
        return buf[:used].copy()

    def interpolate(self, t):
        """dopri_interpolate (R/dopri.R:577-642) for every trajectory of the batch, from the
        history rings in device memory: [n_traj, n, len(t)]."""
        t = _f64(t, "t").ravel()
        y = np.empty((self.n_traj, self.n, t.shape[0]))
        self._check(self._lib.dde_dopri_batch_interpolate(
            self._h, t.ctypes.data_as(c_double_p), t.shape[0], y.ctypes.data_as(c_double_p)))
        return y

    def device_ptr(self, which):
        return self._lib.dde_dopri_batch_device_ptr(self._h, int(which))

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dde_dopri_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DopriShards:
    """One batch resident on several GPUs of the box (dde_dopri_shards_*): a DopriBatch per
    device over contiguous blocks of trajectories, allocated once; run_download_into
    integrates all shards concurrently and streams every shard's results into its slice
    of the caller's (pinned) host arrays.  devices: None = all visible devices."""

    def __init__(self, func, n, n_traj, *, devices=None, n_parms=0, n_out=0, method="dopri5",
                 rtol=1e-6, atol=1e-6, step_size_min=0.0, step_size_max=float("inf"),
                 step_size_initial=0.0, step_max_n=100000, step_size_min_allow=False,
                 stiff_check=0, n_history=0, grow_history=False, return_initial=True):
        self._lib = _lib.load()
        self.ctl = _control(method, rtol, atol, step_size_min, step_size_max, step_size_initial,
                            step_max_n, step_size_min_allow, stiff_check, n_history, grow_history,
                            return_initial)
        self.n, self.n_traj, self.n_parms, self.n_out = int(n), int(n_traj), int(n_parms), int(n_out)
        dev = None if devices is None else np.ascontiguousarray(devices, dtype=np.int32).ravel()
        self._h = self._lib.dde_dopri_shards_alloc(
            func.encode(), C.byref(self.ctl), self.n, self.n_out, self.n_parms, self.n_traj,
            None if dev is None else dev.ctypes.data_as(_lib.c_int32_p),
            0 if dev is None else dev.shape[0])
        if not self._h:
            raise DdeError(_lib.last_error())
        self.n_shards = int(self._lib.dde_dopri_shards_count(self._h))

    def _check(self, rc):
        if rc != 0:
            raise DdeError(_lib.last_error())

    def upload(self, y0=None, parms=None, shared_parms=False):
        stride = 0 if shared_parms else self.n_parms
        self._check(self._lib.dde_dopri_shards_upload(self._h, _ptr(y0), _ptr(parms), stride))

    def set_times(self, times, tcrit=None):
        times = _f64(times, "times").ravel()
        tc = None if tcrit is None else _f64(tcrit, "tcrit").ravel()
        self._check(self._lib.dde_dopri_shards_set_times(
            self._h, times.ctypes.data_as(c_double_p), times.shape[0],
            None if tc is None else tc.ctypes.data_as(c_double_p), 0 if tc is None else tc.shape[0]))
        self.times = times

    def set_events(self, events):
        ev = np.ascontiguousarray(events, dtype=np.int32).ravel()
        self._check(self._lib.dde_dopri_shards_set_events(
            self._h, ev.ctypes.data_as(_lib.c_int32_p), ev.shape[0]))

    def order_by_last_run(self, enable=True):
        """every shard's hand-out order from its own last run (False: back to index order)"""
        self._check(self._lib.dde_dopri_shards_order_by_last_run(self._h, 1 if enable else 0))

    def run_download_into(self, y_out=None, out=None, stats=None, code=None, t_last=None,
                          step_size=None):
        self._check(self._lib.dde_dopri_shards_run_download(
            self._h, _ptr(y_out), _ptr(out), _ptr(stats), _ptr(code), _ptr(t_last),
            _ptr(step_size)))

    def last_ms(self):
        """device time of the slowest shard's last run"""
        return self._lib.dde_dopri_shards_last_ms(self._h)

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dde_dopri_shards_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DifeqBatch:
    def __init__(self, target, n, n_rep, *, n_parms=0, n_out=0, n_history=0, restartable=False):
        self._lib = _lib.load()
        self.n, self.n_rep, self.n_parms, self.n_out = int(n), int(n_rep), int(n_parms), int(n_out)
        self.n_history = int(n_history)
        self._h = self._lib.dde_difeq_batch_alloc(target.encode(), self.n, self.n_out, self.n_parms,
                                                  self.n_rep, self.n_history)
        if not self._h:
            raise DdeError(_lib.last_error())
        if restartable:
            self._lib.dde_difeq_batch_restartable(self._h, 1)

    def _check(self, rc):
        if rc != 0:
            raise DdeError(_lib.last_error())

    def upload(self, y0=None, parms=None, shared_parms=False):
        stride = 0 if shared_parms else self.n_parms
        self._check(self._lib.dde_difeq_batch_upload(self._h, _ptr(y0), _ptr(parms), stride))

    def set_steps(self, steps, return_initial=True):
        st = _steps(steps)
        self._check(self._lib.dde_difeq_batch_set_steps(self._h, st.ctypes.data_as(c_uint64_p),
                                                        st.shape[0], int(bool(return_initial))))
        self.steps = st
        self.nt = st.shape[0] if return_initial else st.shape[0] - 1

    def run(self):
        self._check(self._lib.dde_difeq_batch_run(self._h))

    def continue_(self, steps, y=None, return_initial=True):
        """difeq_continue (src/r_difeq.c:114-165): iterate on from the final step."""
        st = _steps(steps)
        if y is not None and not hasattr(y, "data_ptr"):
            y = np.ascontiguousarray(y, dtype=np.float64)
            if y.size != self.n * self.n_rep:
                raise DdeError("Incorrect size 'y' on simulation restart")
        self._check(self._lib.dde_difeq_batch_continue(
            self._h, _ptr(y), st.ctypes.data_as(c_uint64_p), st.shape[0],
            int(bool(return_initial))))
        self.steps = st
        self.nt = st.shape[0] if return_initial else st.shape[0] - 1

    def copy(self):
        """difeq_data_copy (src/difeq.c:49-72)."""
        h = self._lib.dde_difeq_batch_copy(self._h)
        if not h:
            raise DdeError(_lib.last_error())
        other = object.__new__(DifeqBatch)
        other.__dict__.update(self.__dict__)
        other._h = h
        return other

    def history(self, j):
        """The `history` attribute of replicate j: [entries, 1 + n + n_out] = step, y, out
        (the out columns as the reference leaves them, include/dde_b200.h), oldest first.
        The run must have been restartable."""
        hl = 1 + self.n + self.n_out
        cap = max(self.n_history, 1)
        buf = np.empty((cap, hl))
        used = self._lib.dde_difeq_batch_history(self._h, int(j), buf.ctypes.data_as(c_double_p),
                                                 cap)
        if used < 0:
            raise DdeError(_lib.last_error())
        return buf[:used].copy()

    def sync(self):
        self._check(self._lib.dde_difeq_batch_sync(self._h))

    def last_ms(self):
        return self._lib.dde_difeq_batch_last_ms(self._h)

    def download_into(self, y_out=None, out=None, code=None, y_final=None):
        self._check(self._lib.dde_difeq_batch_download(self._h, _ptr(y_out), _ptr(out), _ptr(code),
                                                       _ptr(y_final)))

    def download(self):
        y_out = np.empty((self.n_rep, self.nt, self.n))
        out = np.empty((self.n_rep, self.nt, self.n_out)) if self.n_out > 0 else None
        code = np.empty((self.n_rep,), dtype=np.int32)
        y_final = np.empty((self.n_rep, self.n))
        self.download_into(y_out, out, code, y_final)
        return y_out, out, code, y_final

    def close(self):
        if getattr(self, "_h", None):
            self._lib.dde_difeq_batch_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
